from .capture import Frame, VideoSource
from .fps_meter import FPSMeter

__all__ = ["VideoSource", "Frame", "FPSMeter"]
