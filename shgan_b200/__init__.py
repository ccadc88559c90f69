"""shgan_b200: B200-native (sm_100a) implementation of the SH-GAN generator-forward hot path behind the
reference's model_zoo module / operator API.  See DESIGN.md and include/shgan_b200.h."""
__version__ = '0.1.0'
