"""Opaque per-process / per-GPU context of libmceik_b200 (device, stream, workspaces, tables)."""
import ctypes as C

from . import _lib


class Context:
    """Owns one ``mceik_ctx``.  ``device`` < 0 means the current CUDA device; ``stream`` may be a
    raw ``cudaStream_t`` integer (e.g. ``torch.cuda.current_stream().cuda_stream``)."""

    def __init__(self, device=-1, stream=None):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.mceik_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise _lib.MceikError(f"mceik_ctx_create failed (rc={rc}): {_lib.last_error()}")
        self.handle = h

    def set_tuning(self, key, value):
        """Development switch of the sweep / search kernels (``mceik_fsm_set_tuning``); returns the context."""
        _lib.check(self._lib.mceik_fsm_set_tuning(self.handle, key.encode(), int(value)), "mceik_fsm_set_tuning")
        return self

    @staticmethod
    def comm_unique_id():
        """``mceik_comm_unique_id``: the 128-byte NCCL id rank 0 hands to the other ranks (bytes)."""
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().mceik_comm_unique_id(buf), "mceik_comm_unique_id")
        return buf.raw

    def comm_init(self, world, rank, unique_id):
        """``mceik_comm_init`` (collective): join the ranks that share the sources of a sharded solve."""
        assert len(unique_id) == 128
        _lib.check(self._lib.mceik_comm_init(self.handle, int(world), int(rank), C.c_char_p(bytes(unique_id))), "mceik_comm_init")
        return self

    def tables_alloc_replicated(self, rows, ldtab):
        """``mceik_tables_alloc_replicated`` (collective): fp32 [rows, ldtab] buffer of this rank that the other ranks
        put their tables into over NVLink; returned as a torch tensor view (the library owns the memory)."""
        import torch
        p = C.c_void_p()
        _lib.check(self._lib.mceik_tables_alloc_replicated(self.handle, int(rows), int(ldtab), C.byref(p)),
                   "mceik_tables_alloc_replicated")

        class _View:  # zero-copy: torch reads __cuda_array_interface__
            __cuda_array_interface__ = {"shape": (int(rows), int(ldtab)), "typestr": "<f4", "data": (p.value, False), "version": 2}
        return torch.as_tensor(_View(), device=torch.device("cuda", torch.cuda.current_device()))

    def tables_free_replicated(self):
        _lib.check(self._lib.mceik_tables_free_replicated(self.handle), "mceik_tables_free_replicated")

    def comm_destroy(self):
        _lib.check(self._lib.mceik_comm_destroy(self.handle), "mceik_comm_destroy")

    def synchronize(self):
        _lib.check(self._lib.mceik_ctx_synchronize(self.handle), "mceik_ctx_synchronize")

    def close(self):
        if getattr(self, "handle", None):
            self._lib.mceik_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
