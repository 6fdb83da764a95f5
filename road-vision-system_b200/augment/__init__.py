"""GPU fog synthesis -- the input generator next to the hot path."""
from .fog import FOG_PRESETS, EnhancedFogSynthesizer, process_folder

__all__ = ["EnhancedFogSynthesizer", "FOG_PRESETS", "process_folder"]
