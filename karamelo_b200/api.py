"""ctypes bindings for include/kml.h + include/kml_host.h.

``Engine`` drives a simulation through the host driver exactly as a user of the
reference would: feed it Karamelo script lines / files, then read state back.
Field ids mirror the enums of ``include/kml.h``.
"""
import ctypes as C
import os

import numpy as np


class KmlError(RuntimeError):
    pass


def lib_dir():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")


class P:  # particle fields (kml.h KML_P_*): id, numpy dtype, trailing shape
    PTAG = (0, np.int64, ())
    X = (1, np.float64, (3,))
    X0 = (2, np.float64, (3,))
    V = (3, np.float64, (3,))
    V_UPDATE = (4, np.float64, (3,))
    A = (5, np.float64, (3,))
    MBP = (6, np.float64, (3,))
    F = (7, np.float64, (3,))
    SIGMA = (8, np.float64, (3, 3))
    STRAIN_EL = (9, np.float64, (3, 3))
    VOL0PK1 = (10, np.float64, (3, 3))
    FDEF = (11, np.float64, (3, 3))
    R = (12, np.float64, (3, 3))
    J = (13, np.float64, ())
    VOL0 = (14, np.float64, ())
    VOL = (15, np.float64, ())
    RHO0 = (16, np.float64, ())
    RHO = (17, np.float64, ())
    MASS = (18, np.float64, ())
    EFF_PLASTIC_STRAIN = (19, np.float64, ())
    EFF_PLASTIC_STRAIN_RATE = (20, np.float64, ())
    DAMAGE = (21, np.float64, ())
    DAMAGE_INIT = (22, np.float64, ())
    IENERGY = (23, np.float64, ())
    MASK = (24, np.int32, ())
    T = (25, np.float64, ())
    GAMMA = (26, np.float64, ())
    Q = (27, np.float64, (3,))


class N:  # node fields (kml.h KML_N_*)
    X0 = (0, np.float64, (3,))
    X = (1, np.float64, (3,))
    V = (2, np.float64, (3,))
    V_UPDATE = (3, np.float64, (3,))
    MB = (4, np.float64, (3,))
    F = (5, np.float64, (3,))
    MASS = (6, np.float64, ())
    MASK = (7, np.int32, ())
    NTYPE = (8, np.int32, (3,))
    RIGID = (9, np.int32, ())
    T = (10, np.float64, ())
    T_UPDATE = (11, np.float64, ())
    QEXT = (12, np.float64, ())
    QINT = (13, np.float64, ())


STAGES = ["rebin", "p2g", "grid", "g2p", "v2g", "stress", "contact", "other", "halo", "migrate", "dt"]


def load_host_library(path=None):
    """Load libkml_host.so (and through it the engine library it was linked with).

    ``path=None`` loads the product build in ``karamelo_b200/lib``; it raises if the CUDA
    libraries are missing - there is no CPU fallback in the product.
    """
    if path is None:
        path = os.path.join(lib_dir(), "libkml_host.so")
        if not os.path.exists(path) or not os.path.exists(os.path.join(lib_dir(), "libkml.so")):
            raise KmlError(
                "karamelo_b200: the CUDA engine is not built (%s missing). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C karamelo_b200`. There is no CPU fallback." % path)
    lib = C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    PD, PI, PL = C.POINTER(dbl), C.POINTER(i32), C.POINTER(i64)
    sig = {
        "kmlh_last_error": (C.c_char_p, []),
        "kmlh_create": (i32, [C.POINTER(vp)]),
        "kmlh_destroy": (i32, [vp]),
        "kmlh_set_quiet": (i32, [vp, i32]),
        "kmlh_set_device": (i32, [vp, i32]),
        "kmlh_run_file": (i32, [vp, C.c_char_p]),
        "kmlh_run_line": (i32, [vp, C.c_char_p]),
        "kmlh_get_var": (i32, [vp, C.c_char_p, PD]),
        "kmlh_nsolids": (i32, [vp, PI]),
        "kmlh_solid_info": (i32, [vp, i32, PL, PI, PI, PI]),
        "kmlh_state": (i32, [vp, PL, PD, PD]),
        "kmlh_ctx": (vp, [vp]),
        "kmlh_apply_initial_fixes": (i32, [vp]),
        "kmlh_eval": (i32, [vp, C.c_char_p, PD]),
        "kmlh_set_particle_var": (i32, [vp, C.c_char_p, dbl]),
        "kmlh_compile_expr": (i32, [vp, C.c_char_p, PI, PI, PD]),
        "kmlh_set_ranks": (i32, [vp, i32, i32, vp]),
        "kmlh_slab_info": (i32, [vp, i32, PL]),
        "kml_comm_unique_id": (i32, [vp]),
        # engine ABI, resolved through the host library's dependency
        "kml_last_error": (C.c_char_p, []),
        "kml_backend": (C.c_char_p, []),
        "kml_solid_download": (i32, [vp, i32, i32, vp]),
        "kml_solid_upload": (i32, [vp, i32, i32, vp]),
        "kml_grid_download": (i32, [vp, i32, i32, vp]),
        "kml_grid_upload": (i32, [vp, i32, i32, vp]),
        "kml_grid_nnodes": (i32, [vp, i32, PL]),
        "kml_solid_np": (i32, [vp, i32, PL]),
        "kml_solid_generation": (i32, [vp, i32, C.POINTER(C.c_uint64)]),
        "kml_synchronize": (i32, [vp]),
        "kml_profile": (i32, [vp, i32]),
        "kml_stage_times": (i32, [vp, PD, PL, i32]),
        "kml_stage_host_times": (i32, [vp, PD, i32]),
        "kml_error_flags": (i32, [vp, C.POINTER(C.c_uint)]),
        "kml_get_dt": (i32, [vp, PD]),
        "kml_timer_start": (i32, [vp]),
        "kml_timer_stop": (i32, [vp, PD]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    return lib


class Engine:
    """One simulation: a script interpreter + device state."""

    def __init__(self, lib=None, quiet=True, device=0, rank=0, nranks=1, nccl_id=None):
        """``rank``/``nranks``: this process drives slab ``rank`` of ``nranks`` (one process per GPU);
        ``nccl_id`` is the 128-byte id of ``nccl_unique_id()`` on rank 0, shared with every rank."""
        self.lib = lib if lib is not None and not isinstance(lib, str) else load_host_library(lib)
        h = C.c_void_p()
        self._ck(self.lib.kmlh_create(C.byref(h)))
        self.h = h
        self.lib.kmlh_set_quiet(h, 1 if quiet else 0)
        self.lib.kmlh_set_device(h, device)
        self.rank, self.nranks = rank, nranks
        if nranks > 1:
            buf = C.create_string_buffer(bytes(nccl_id), 128) if nccl_id is not None else None
            self._ck(self.lib.kmlh_set_ranks(h, rank, nranks, C.cast(buf, C.c_void_p) if buf is not None else None))

    def nccl_unique_id(self):
        buf = C.create_string_buffer(128)
        self._ckk(self.lib.kml_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf.raw)

    def slab_info(self, i=0):
        info = (C.c_int64 * 8)()
        self._ck(self.lib.kmlh_slab_info(self.h, i, info))
        keys = ["base_lo", "base_hi", "goff", "nx_local", "own_lo", "own_hi", "np_global", "tag_offset"]
        return dict(zip(keys, [int(v) for v in info]))

    # -- errors -------------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise KmlError((self.lib.kmlh_last_error() or b"").decode(errors="replace"))

    def _ckk(self, rc):
        if rc:
            raise KmlError((self.lib.kml_last_error() or b"").decode(errors="replace"))

    @property
    def backend(self):
        return self.lib.kml_backend().decode()

    def close(self):
        if self.h:
            self.lib.kmlh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- script -------------------------------------------------------------------------
    def line(self, text):
        self._ck(self.lib.kmlh_run_line(self.h, text.encode()))

    def var_eval(self, expr):
        v = C.c_double(0)
        self._ck(self.lib.kmlh_eval(self.h, expr.encode(), C.byref(v)))
        return v.value

    def apply_initial_fixes(self):
        self._ck(self.lib.kmlh_apply_initial_fixes(self.h))

    def script(self, text):
        for ln in text.splitlines():
            if ln.split("#")[0].strip() == "quit":
                break
            self.line(ln)

    def file(self, path):
        self._ck(self.lib.kmlh_run_file(self.h, str(path).encode()))

    def var(self, name):
        v = C.c_double()
        self._ck(self.lib.kmlh_get_var(self.h, name.encode(), C.byref(v)))
        return v.value

    # -- state --------------------------------------------------------------------------
    @property
    def ctx(self):
        return self.lib.kmlh_ctx(self.h)

    def nsolids(self):
        n = C.c_int()
        self._ck(self.lib.kmlh_nsolids(self.h, C.byref(n)))
        return n.value

    def solid_info(self, i):
        np_, sid, gid = C.c_int64(), C.c_int(), C.c_int()
        n = (C.c_int * 3)()
        self._ck(self.lib.kmlh_solid_info(self.h, i, C.byref(np_), C.byref(sid), C.byref(gid), n))
        cur = C.c_int64()
        self._ckk(self.lib.kml_solid_np(self.ctx, sid.value, C.byref(cur)))
        return {"np": cur.value, "solid": sid.value, "grid": gid.value, "n": tuple(n)}

    def solid_generation(self, i):
        """Bumped whenever the solid's particle set or storage order changed (migration, physical permute, delete_particles)."""
        g = C.c_uint64()
        self._ckk(self.lib.kml_solid_generation(self.ctx, self.solid_info(i)["solid"], C.byref(g)))
        return g.value

    def state(self):
        nt, t, dt = C.c_int64(), C.c_double(), C.c_double()
        self._ck(self.lib.kmlh_state(self.h, C.byref(nt), C.byref(t), C.byref(dt)))
        return {"ntimestep": nt.value, "time": t.value, "dt": dt.value}

    def download(self, isolid, field):
        info = self.solid_info(isolid)
        fid, dt, shp = field
        out = np.empty((info["np"],) + shp, dtype=dt)
        self._ckk(self.lib.kml_solid_download(self.ctx, info["solid"], fid, out.ctypes.data_as(C.c_void_p)))
        return out

    def upload(self, isolid, field, arr):
        info = self.solid_info(isolid)
        fid, dt, shp = field
        a = np.ascontiguousarray(arr, dtype=dt)
        assert a.shape == (info["np"],) + shp, (a.shape, info["np"], shp)
        self._ckk(self.lib.kml_solid_upload(self.ctx, info["solid"], fid, a.ctypes.data_as(C.c_void_p)))

    def grid_download(self, isolid, field):
        info = self.solid_info(isolid)
        nn = C.c_int64()
        self._ckk(self.lib.kml_grid_nnodes(self.ctx, info["grid"], C.byref(nn)))
        fid, dt, shp = field
        out = np.empty((nn.value,) + shp, dtype=dt)
        self._ckk(self.lib.kml_grid_download(self.ctx, info["grid"], fid, out.ctypes.data_as(C.c_void_p)))
        return out

    def snapshot(self, fields=("PTAG", "X", "V", "SIGMA", "STRAIN_EL", "FDEF", "EFF_PLASTIC_STRAIN", "EFF_PLASTIC_STRAIN_RATE", "DAMAGE", "DAMAGE_INIT", "VOL", "MASS")):
        """Per-solid dict of arrays sorted by particle tag (stable across re-binning)."""
        snaps = []
        for i in range(self.nsolids()):
            d = {f: self.download(i, getattr(P, f)) for f in fields}
            order = np.argsort(d["PTAG"], kind="stable")
            snaps.append({k: v[order] for k, v in d.items()})
        return snaps

    # -- measurement --------------------------------------------------------------------
    def synchronize(self):
        self._ckk(self.lib.kml_synchronize(self.ctx))

    def profile(self, enable=True):
        self._ckk(self.lib.kml_profile(self.ctx, 1 if enable else 0))

    def stage_times(self, reset=True):
        ms = (C.c_double * len(STAGES))()
        ln = (C.c_int64 * len(STAGES))()
        self._ckk(self.lib.kml_stage_times(self.ctx, ms, ln, 1 if reset else 0))
        return {s: (ms[i], ln[i]) for i, s in enumerate(STAGES)}

    def stage_host_times(self, reset=True):
        ms = (C.c_double * len(STAGES))()
        self._ckk(self.lib.kml_stage_host_times(self.ctx, ms, 1 if reset else 0))
        return {s: ms[i] for i, s in enumerate(STAGES)}

    def timer_start(self):
        self._ckk(self.lib.kml_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_double()
        self._ckk(self.lib.kml_timer_stop(self.ctx, C.byref(ms)))
        return ms.value

    def error_flags(self):
        f = C.c_uint()
        self._ckk(self.lib.kml_error_flags(self.ctx, C.byref(f)))
        return f.value
