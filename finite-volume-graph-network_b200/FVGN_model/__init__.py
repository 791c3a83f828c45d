from .FVGN import FVGN  # noqa: F401
from .EncoderProcesserDecoder import (EncoderProcesserDecoder, Encoder, GnBlock, Decoder, Intergrator, build_mlp)  # noqa: F401
from .GN_blocks import CellBlock, EdgeBlock  # noqa: F401
