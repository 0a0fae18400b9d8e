from .discriminators import ProgressiveDiscriminator, ProgressiveDiscriminator_inputCat, ProgressiveEncoderDiscriminator  # noqa: F401
