from .deform_attn import (CustomMSDeformableAttention, MSDeformableAttention3D,  # noqa: F401
                          PredictionMSDeformableAttention, SpatialCrossAttention, TemporalSelfAttention)
from .latent_rendering import LatentRendering, latent_render_core  # noqa: F401
