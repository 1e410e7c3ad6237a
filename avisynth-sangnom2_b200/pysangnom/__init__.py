"""Host-side Python helpers for the B200 SangNom2 path: ctypes bindings of the C-ABI library
(`cuda.py`), the fake AviSynth host driver used by the plugin tests (`fakehost.py`) and the seeded
synthetic clip generators (`clips.py`). The compute lives in ../csrc (CUDA, sm_100a only)."""
