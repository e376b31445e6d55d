"""Start the CUDA context of the first device while the interpreter is still importing numpy and friends.

Creating the primary context on a 180 GB device takes most of a second and so does importing the numeric stack; the
drop-in scripts that always need the GPU call warm() before their heavy imports, so that the two overlap (the library
call releases the GIL).  The context created here is the device's PRIMARY context: the handle is destroyed again at
once, the context stays, and the Engine the script creates later finds it ready.  Every failure is swallowed -- the
Engine reports it properly.  Only ctypes / os / threading are imported here; never used by scripts that may fork
worker processes afterwards (a forked child must not inherit a live context)."""
import os
import threading


def first_device():
    v = os.environ.get("LONGSOM_GPUS", "").strip()
    if v and "," in v:
        try:
            return int(v.split(",")[0])
        except ValueError:
            return 0
    return 0


def warm(device=None):
    if os.environ.get("LONGSOM_EARLY_CUDA", "1") == "0":
        return None
    dev = first_device() if device is None else int(device)

    def run():
        try:
            import ctypes as C
            from . import _lib
            lib = _lib.load()
            ctx = C.c_void_p()
            if lib.ls_ctx_create(dev, C.byref(ctx)) == _lib.LS_OK:
                lib.ls_ctx_destroy(ctx)
        except Exception:
            pass
    t = threading.Thread(target=run, daemon=True)
    t.start()
    return t
