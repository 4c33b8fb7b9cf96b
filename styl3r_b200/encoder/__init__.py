from .encoder import (BackboneCrocoCfg, EncoderNoPoSplatMultiTokenStyle, EncoderNoPoSplatTokenStyleCfg,  # noqa: F401
                      GaussianAdapterCfg, Gaussians, OpacityMappingCfg, TokenStylizerCfg, get_encoder)
