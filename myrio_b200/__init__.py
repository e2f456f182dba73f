"""myrio_b200 -- B200-native (sm_100a CUDA) implementation of Myrio's k-mer
counting / leaf scoring / read clustering hot path behind a C ABI
(include/myrio_b200.h).  The CUDA library is loaded on first use and there is no
CPU fallback: a missing libmyrio_b200.so raises."""
__version__ = "0.1.0"
