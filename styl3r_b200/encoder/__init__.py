from .encoder import (BackboneCrocoCfg, EncoderNoPoSplatMultiTokenStyle, EncoderNoPoSplatTokenStyleCfg,  # noqa: F401
                      GaussianAdapterCfg, Gaussians, GraphedEncoder, OpacityMappingCfg, TokenStylizerCfg, get_encoder)
