from .capture import Frame, SyntheticReader, VideoSource
from .feeder import Batch, BatchFeeder
from .fps_meter import FPSMeter

__all__ = ["VideoSource", "Frame", "FPSMeter", "SyntheticReader", "BatchFeeder", "Batch"]
