"""Host-side Python helpers for the B200 SangNom2 path: ctypes bindings of the C-ABI library
(`cuda.py`), the planar format table (`formats.py`), frame-range sharding (`shard.py`) and the seeded
synthetic clip generators (`clips.py`). The compute lives in ../csrc (CUDA, sm_100a only)."""
