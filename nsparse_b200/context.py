"""Per-GPU context: owns the native nsp_context (stream, workspace arena, plan state)."""
from __future__ import annotations

import ctypes as C

from . import _lib


class Context:
    def __init__(self, device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise _lib.NsparseError("no CUDA device visible: nsparse_b200 runs on B200 (sm_100a) only")
        self.lib = _lib.load()
        self.device = device
        h = C.c_void_p()
        rc = self.lib.nsp_create(C.byref(h), device)
        if rc != 0:
            raise _lib.NsparseError(f"nsp_create(device={device}) failed with {rc}")
        self.handle = h
        self.use_torch_stream()

    def use_torch_stream(self):
        """Run on torch's current stream of this device, so torch.cuda.Event timing sees us."""
        import torch

        s = torch.cuda.current_stream(self.device).cuda_stream
        self.check(self.lib.nsp_set_stream(self.handle, C.c_void_p(s)))

    def check(self, rc: int):
        if rc != 0:
            msg = self.lib.nsp_last_error(self.handle)
            raise _lib.NsparseError(f"nsparse_b200 error {rc}: {msg.decode() if msg else ''}")

    def set_option(self, name: str, value: int):
        self.check(self.lib.nsp_set_option(self.handle, name.encode(), value))

    def sync(self):
        self.check(self.lib.nsp_sync(self.handle))

    @property
    def launches(self) -> int:
        return int(self.lib.nsp_launch_count(self.handle))

    def profile(self, on: bool = True):
        self.set_option("profile", 1 if on else 0)

    def profile_dump(self):
        """[(kernel, ms, rows, intermediate products, A entries, C entries)] of the launches since the
        last dump (C entries is 0 for the symbolic kernels)."""
        buf = C.create_string_buffer(1 << 16)
        self.check(self.lib.nsp_profile_dump(self.handle, buf, len(buf)))
        out = []
        for line in buf.value.decode().splitlines():
            n, ms, rows, ip, alen, nout = line.split()
            out.append((n, float(ms), int(rows), int(ip), int(alen), int(nout)))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.nsp_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = {}


def default_context(device: int = 0) -> Context:
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
