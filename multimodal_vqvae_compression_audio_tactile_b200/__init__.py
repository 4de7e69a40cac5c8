"""B200-native (sm_100a) encode -> quantize -> decode path of the audio/vibrotactile VQ-VAE codec
of aymenboudhina/Multimodal_VQVAE_compression_audio_tactile, behind the reference's own
nn.Module interface.  See DESIGN.md / INTEGRATION.md."""
from ._lib import B2CError, LIB_PATH, load  # noqa: F401
from .modules import (AR_CHUNK_TOK, CODE_DIM, DAC, CrossPredictor, Decoder, Encoder, PosEnc1D, ProposedEval,  # noqa: F401
                      ResidualVectorQuantize, ResidualVQEMA, TokenNorm, build_proposed)
from .ops import nearest_code  # noqa: F401
from . import metrics  # noqa: F401
from .plc import (AllPredPLC, build_plc, make_category_token_loss_mask_for_category,  # noqa: F401
                  make_token_loss_mask)
from .bitstream import bits_per_index, estimated_kbps, pack_indices, packed_bytes, unpack_indices  # noqa: F401

#: contraction arithmetic bench.py / smoke() use by default (see DESIGN.md "Precision")
DEFAULT_PRECISION = "tc"
