"""rabbit-transcoding_b200 — B200-native hot path of RABBIT's V-PCC transcode loop.

Only what the path needs: `csrc/` (hand-written sm_100a kernels + the C ABI of include/rabbit_b200.h),
`abi.py` (ctypes mirror), `codec.py` / `metrics.py` (host-side mirror of the reference's PCCCodec / PCCMetrics
entry points for this path), `synthetic.py` (seeded decoded-GOF generator), `dist.py` (frame sharding).
The directory name carries a hyphen; import it as `rabbit_transcoding_b200` (see the loader at the repo root).
"""
from . import abi, codec, dist, metrics, synthetic  # noqa: F401
