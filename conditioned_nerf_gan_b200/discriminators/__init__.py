from .discriminators import ProgressiveDiscriminator  # noqa: F401
