"""b200vgan -- B200-native BigVGAN2 speech-code decoder (drop-in for indextts.BigVGAN.models.BigVGAN).

Importing this package never touches the oracle and never falls back to a CPU path."""
from . import lib, sched, synth  # noqa: F401
from .lib import BvgError, MODE_BF16, MODE_F16, MODE_FP32  # noqa: F401


def __getattr__(name):
    if name in ("BigVGAN", "Generator"):
        from .model import BigVGAN
        return BigVGAN
    if name == "MelSpectrogramFeatures":
        from .features import MelSpectrogramFeatures
        return MelSpectrogramFeatures
    raise AttributeError(name)
