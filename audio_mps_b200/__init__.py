"""audio_mps_b200 -- B200-native (sm_100a) implementation of the AudioMPS continuous-MPS scan
behind the reference's model-level API.  (The directory is spelled with an underscore because a
hyphen is not importable in Python; DESIGN.md, "Layout".)"""
from .hparams import HParams, default_hparams
from .model import CMPS, PsiCMPS, RhoCMPS
from .data import DeviceBatchPrefetcher, get_audio, damped_sine, random_raw_params, sample_noise

__all__ = ["HParams", "default_hparams", "CMPS", "PsiCMPS", "RhoCMPS", "get_audio", "damped_sine",
           "DeviceBatchPrefetcher", "random_raw_params", "sample_noise"]
