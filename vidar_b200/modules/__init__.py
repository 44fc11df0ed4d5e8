from .deform_attn import (MSDeformableAttention3D, PredictionMSDeformableAttention,  # noqa: F401
                          SpatialCrossAttention, TemporalSelfAttention)
from .latent_rendering import LatentRendering, latent_render_core  # noqa: F401
