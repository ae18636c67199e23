"""Import-compatible alias of the reference module name (``import AXCTDprocessor``)."""
from axctdprocessor_b200.AXCTDprocessor import *  # noqa: F401,F403
from axctdprocessor_b200.AXCTDprocessor import AXCTD_Processor, readAXCTDwavfile  # noqa: F401
