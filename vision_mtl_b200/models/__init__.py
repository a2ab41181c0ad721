from .basic_model import BasicMTLModel  # noqa: F401
from .cross_stitch_model import CrossStitchLayer, CSNet  # noqa: F401
from .mtan_model import (  # noqa: F401
    AttentionModuleDecoder,
    AttentionModuleEncoder,
    MTANDown,
    MTANMiniUnet,
    MTANUp,
)
