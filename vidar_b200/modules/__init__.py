from .deform_attn import (MSDeformableAttention3D, PredictionMSDeformableAttention,  # noqa: F401
                          SpatialCrossAttention, TemporalSelfAttention)
