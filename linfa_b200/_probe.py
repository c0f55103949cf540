"""ctypes access to liblkm_probe.so: the product sources plus the tcgen05 probes and micro-benchmarks
(csrc/lkm_probe.cuh).  Development / evidence only — nothing in linfa_b200's product path imports this."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblkm_probe.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build it with `python linfa_b200/build.py`")
lib = C.CDLL(LIB_PATH)
lib.lkm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
lib.lkm_destroy.argtypes = [C.c_void_p]
lib.lkm_last_error.argtypes = [C.c_void_p]
lib.lkm_last_error.restype = C.c_char_p


class ProbeContext:
    """A context of the probe library (its lkm_ctx is private to that library)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        st = lib.lkm_create(C.byref(self.h), device)
        if st != 0:
            raise RuntimeError(f"lkm_create (probe library) failed: {st}")

    def error(self):
        return lib.lkm_last_error(self.h).decode()

    def close(self):
        if self.h:
            lib.lkm_destroy(self.h)
            self.h = None
