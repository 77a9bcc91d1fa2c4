from .model import ResidualUNet3D, UNet3D, create_feature_maps  # noqa: F401
