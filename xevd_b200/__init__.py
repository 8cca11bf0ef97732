"""xevd_b200 -- B200-native (sm_100a CUDA) reconstruction path of the XEVD MPEG-5 EVC decoder."""
__version__ = "0.1.0"
