"""mythos_b200 -- B200-native (sm_100a) kernels for mythos' oxDNA-family energy / force / theta-gradient path.

Host code is Python mirroring the reference's EnergyFunction / simulator_init / DiffTRe interfaces; all
arithmetic on the hot path runs in hand-written CUDA behind the C-ABI of ``include/mythos_b200.h``.
"""

__version__ = "0.1.0"
