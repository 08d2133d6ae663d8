"""transtacos-retunegan_b200 -- B200-native spectral front/back end of TransTacoS / RetuneGAN.

Hand-written sm_100a CUDA kernels (libspectral_b200.so, C ABI in include/spectral_b200.h) behind the
reference's own Python function API:

    transtacos_audio  <->  transtacos/audio.py   (get_specs, inv_spec, preemphasis, ...)
    retunegan_audio   <->  retunegan/audio.py    (get_mag, get_mel, mag_to_mel, inv_mag, get_stft_torch)
    loss              <->  retunegan/models/loss.py::multi_stft_loss

The directory name contains a hyphen; import it as ``import transtacos_retunegan_b200`` (alias module at the
repository root) or ``importlib.import_module("transtacos-retunegan_b200")``.
"""
from . import config, _lib, core, sharding            # noqa: F401
from . import transtacos_audio, retunegan_audio, loss, preprocess, retunegan_data  # noqa: F401
from .config import SpectralConfig, TRANSTACOS, RETUNEGAN, PI   # noqa: F401
from .loss import multi_stft_loss, envelope_loss, dynamic_loss   # noqa: F401

# keithito-style aliases named in BASELINE.json's north_star
spectrogram = lambda y: transtacos_audio.get_specs(y)[0]        # noqa: E731
melspectrogram = lambda y: transtacos_audio.get_specs(y)[1]     # noqa: E731
inv_spectrogram = transtacos_audio.inv_spec

__all__ = ["config", "core", "sharding", "transtacos_audio", "retunegan_audio", "loss", "SpectralConfig",
           "TRANSTACOS", "RETUNEGAN", "PI", "multi_stft_loss", "spectrogram", "melspectrogram", "inv_spectrogram"]
