"""egom2p_b200: B200-native (sm_100a) implementation of the EgoM2P masked multimodal training step behind the
reference's model / adapter API. See DESIGN.md and INTEGRATION.md."""
from .registry import create_model, register_model, register_into_reference  # noqa: F401
from .model import EgoM2P  # noqa: F401
from .modality_info import MODALITY_INFO, MOD4  # noqa: F401
