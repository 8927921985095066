"""ctypes binding of the C ABI in include/ue_gpu.h.

`UeLib(path, prefix)` binds one shared library that exports the ABI under a
given symbol prefix.  The product library is `uedge_b200/csrc/libuegpu.so`
(prefix ``ue_gpu_``); `load_gpu()` loads it and fails loudly if it is missing —
there is no CPU fallback on the product path.  Tests bind the CPU oracle with
the same class (prefix ``ue_ora_``) purely as a checker.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB = os.environ.get("UE_GPU_LIB", os.path.join(_HERE, "csrc", "libuegpu.so"))

_i64 = C.c_int64
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class UeError(RuntimeError):
    """Raised where the reference would call xerrab (com/error.f:1-13)."""


class UeLib:
    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s not found — build it first (python -c 'import __graft_entry__ as g; g.build()')" % path
            )
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.neq = None
        f = self._f
        f("set_int", [C.c_char_p, _i64])
        f("set_real", [C.c_char_p, C.c_double])
        f("set_real_array", [C.c_char_p, _dp, _i64])
        f("set_int_array", [C.c_char_p, _ip, _i64])
        f("init", [])
        f("step_params", [_i64, _dp, _dp, _dp, _dp])
        f("pandf1", [_i64, C.c_double, _dp, _dp])
        f("jac_calc", [_i64, C.c_double, _dp, _dp, _i64, _i64, _i64, _dp, _ip, _ip, _ip])
        f("set_column_range", [_i64, _i64])
        getattr(self.lib, prefix + "last_error").restype = C.c_char_p

    def _f(self, name, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
        return fn

    def _call(self, name, *args):
        rc = getattr(self.lib, self.prefix + name)(*args)
        if rc != 0:
            msg = getattr(self.lib, self.prefix + "last_error")().decode()
            raise UeError("%s%s failed (%d): %s" % (self.prefix, name, rc, msg))

    # ---- static inputs -------------------------------------------------------
    def load_static(self, static):
        for k, v in static["ints"].items():
            self._call("set_int", k.encode(), int(v))
        for k, v in static["reals"].items():
            self._call("set_real", k.encode(), float(v))
        for k, v in static.get("zero_ints", {}).items():
            self._call("set_int", k.encode(), int(v))
        for k, v in static.get("zero_reals", {}).items():
            self._call("set_real", k.encode(), float(v))
        for grp in ("planes", "lines"):
            for k, v in static[grp].items():
                a = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
                self._call("set_real_array", k.encode(), _d(a), a.size)
        for grp in ("iplanes", "ilines"):
            for k, v in static[grp].items():
                a = np.ascontiguousarray(v, dtype=np.int64).reshape(-1)
                self._call("set_int_array", k.encode(), _i(a), a.size)
        self.neq = int(static["ints"]["neq"])
        self.nnzmx_default = None

    def set_real(self, name, v):
        self._call("set_real", name.encode(), float(v))

    def set_int(self, name, v):
        self._call("set_int", name.encode(), int(v))

    def init(self):
        self._call("init")

    def step_params(self, dtuse, ylodt, suscal, sfscal):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (dtuse, ylodt, suscal, sfscal)]
        self._call("step_params", self.neq, *[_d(x) for x in a])

    # ---- hot path ---------------------------------------------------------------
    def set_dt(self, yl):
        """set_dt of the nksol driver: (f0, dtuse)."""
        fn = getattr(self.lib, self.prefix + "set_dt")
        fn.argtypes = [_i64, _dp, _dp, _dp]
        fn.restype = C.c_int
        yl = np.ascontiguousarray(yl, dtype=np.float64)
        f0 = np.zeros(self.neq); dt = np.zeros(self.neq)
        if fn(self.neq, _d(yl), _d(f0), _d(dt)) != 0:
            raise UeError("set_dt failed: %s" % getattr(self.lib, self.prefix + "last_error")().decode())
        return f0, dt

    def pandf1(self, yl, time=0.0, out=None):
        """Pandf1rhs_interface: full residual pandf1(-1,-1,0,neq,time,yl,yldot).  `out` reuses a caller buffer."""
        yl = np.ascontiguousarray(yl, dtype=np.float64)
        assert yl.size == self.neq + 2
        yldot = np.zeros(self.neq) if out is None else out
        assert yldot.dtype == np.float64 and yldot.size >= self.neq and yldot.flags.c_contiguous
        self._call("pandf1", self.neq, float(time), _d(yl), _d(yldot))
        return yldot

    def jac_calc(self, yl, yldot00, ml, mu, nnzmx, t=0.0, out=None):
        """jac_calc_interface: returns (jac, ja, ia) in the reference's 1-based CSR.  `out` = (jac, ja, ia) caller buffers
        (e.g. page-locked ones); yldot00 is then passed as given if it is a float64 array of at least neq entries."""
        yl = np.ascontiguousarray(yl, dtype=np.float64)
        if out is None:
            y0 = np.zeros(self.neq + 2)
            y0[: self.neq] = np.asarray(yldot00, dtype=np.float64)[: self.neq]
            jac = np.zeros(nnzmx); ja = np.zeros(nnzmx, dtype=np.int64); ia = np.zeros(self.neq + 1, dtype=np.int64)
        else:
            jac, ja, ia = out
            y0 = yldot00
            assert y0.dtype == np.float64 and y0.size >= self.neq and jac.size >= nnzmx and ja.size >= nnzmx and ia.size >= self.neq + 1
        nnz = C.c_int64(0)
        self._call("jac_calc", self.neq, float(t), _d(yl), _d(y0), int(ml), int(mu), int(nnzmx), _d(jac), _i(ja), _i(ia),
                   C.byref(nnz))
        n = nnz.value
        return jac[:n].copy(), ja[:n].copy(), ia.copy() if out is not None else ia

    def rhs_jac(self, yl, ml, mu, nnzmx):
        """rhsnk(yl) + jac_calc(yl, yldot00) in one call (what psetnk/sfsetnk issue back to back); product library only."""
        fn = getattr(self.lib, self.prefix + "rhs_jac")
        fn.argtypes = [_i64, _dp, _dp, _i64, _i64, _i64, _dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        fn.restype = C.c_int
        yl = np.ascontiguousarray(yl, dtype=np.float64)
        f = np.zeros(self.neq); jac = np.zeros(nnzmx); ja = np.zeros(nnzmx, dtype=np.int64); ia = np.zeros(self.neq + 1, dtype=np.int64)
        nnz = C.c_int64(0)
        if fn(self.neq, _d(yl), _d(f), int(ml), int(mu), int(nnzmx), _d(jac), _i(ja), _i(ia), C.byref(nnz)) != 0:
            raise UeError("rhs_jac failed: %s" % getattr(self.lib, self.prefix + "last_error")().decode())
        n = nnz.value
        return f, (jac[:n].copy(), ja[:n].copy(), ia)

    def jac_scale(self, su, sf, nnz, isrnorm=1, normtype=0):
        """psetnk's scaling chain on the device-resident Jacobian of the last jac_calc (bbb/oderhs.m:9473-9485); product library only."""
        fn = getattr(self.lib, self.prefix + "jac_scale")
        fn.argtypes = [_i64, _dp, _dp, _i64, _i64, _i64, _dp, _dp]
        fn.restype = C.c_int
        su = np.ascontiguousarray(su, dtype=np.float64); sf = np.ascontiguousarray(sf, dtype=np.float64)
        jac = np.zeros(nnz); fac = np.zeros(self.neq)
        if fn(self.neq, _d(su), _d(sf), int(isrnorm), int(normtype), int(nnz), _d(jac), _d(fac)) != 0:
            raise UeError("jac_scale failed: %s" % getattr(self.lib, self.prefix + "last_error")().decode())
        return jac, fac

    def sfsetnk(self, yl, su, ml, mu):
        """Row scale factors sf and ydt_max0 as sfsetnk computes them (bbb/oderhs.m:9815-9884); product library only."""
        fn = getattr(self.lib, self.prefix + "sfsetnk")
        fn.argtypes = [_i64, _dp, _dp, _i64, _i64, _dp, _dp]
        fn.restype = C.c_int
        yl = np.ascontiguousarray(yl, dtype=np.float64)
        su = np.ascontiguousarray(su, dtype=np.float64)
        sf = np.zeros(self.neq)
        ym = C.c_double(0.0)
        rc = fn(self.neq, _d(yl), _d(su), int(ml), int(mu), _d(sf), C.byref(ym))
        if rc != 0:
            raise UeError("sfsetnk failed (%d): %s" % (rc, getattr(self.lib, self.prefix + "last_error")().decode()))
        return sf, ym.value

    def set_column_range(self, ivmin, ivmax):
        self._call("set_column_range", int(ivmin), int(ivmax))


def load_gpu():
    """The product library.  Raises if it has not been built: no fallback."""
    return UeLib(GPU_LIB, "ue_gpu_")


def split_init(lib, world, rank, dist, torch, transport="p2p"):
    """Bind this rank's library instance to a WORLD-rank column split of the Jacobian (include/ue_gpu.h, multi-GPU block).
    `dist` (torch.distributed) plays the part of the Fortran host's MPI: it only carries the 128-byte NCCL id or the
    64-byte CUDA IPC handles between the processes and the host barrier; the data path is the library's own."""
    import ctypes as C
    if transport == "nccl":
        lib.ue_gpu_comm_unique_id.argtypes = [C.c_char_p]
        lib.ue_gpu_comm_init.argtypes = [C.c_int64, C.c_int64, C.c_char_p]
        idbuf = C.create_string_buffer(128)
        if rank == 0 and lib.ue_gpu_comm_unique_id(idbuf) != 0:
            raise RuntimeError(lib.ue_gpu_last_error().decode())
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        rc = lib.ue_gpu_comm_init(world, rank, bytes(t.cpu().numpy().tobytes()))
    elif transport == "p2p":
        lib.ue_gpu_comm_p2p_handle.argtypes = [C.c_char_p]
        lib.ue_gpu_comm_init_p2p.argtypes = [C.c_int64, C.c_int64, C.c_char_p]
        hbuf = C.create_string_buffer(64)
        if lib.ue_gpu_comm_p2p_handle(hbuf) != 0:
            raise RuntimeError(lib.ue_gpu_last_error().decode())
        mine = torch.frombuffer(bytearray(hbuf.raw), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        rc = lib.ue_gpu_comm_init_p2p(world, rank, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    else:
        raise ValueError("transport: nccl or p2p")
    if rc != 0:
        raise RuntimeError(lib.ue_gpu_last_error().decode())
    torch.cuda.synchronize()
    dist.barrier()


def split_init_gen(lib, world, rank, dist, torch):
    """The same for the general path (ue_gen_comm_init, include/ue_gen.h): NCCL id from rank 0, then the collective init."""
    import ctypes as C
    lib.ue_gpu_comm_unique_id.argtypes = [C.c_char_p]
    lib.ue_gen_comm_init.argtypes = [C.c_int64, C.c_int64, C.c_char_p]
    idbuf = C.create_string_buffer(128)
    if rank == 0 and lib.ue_gpu_comm_unique_id(idbuf) != 0:
        raise RuntimeError(lib.ue_gpu_last_error().decode())
    t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
    dist.broadcast(t, 0)
    if lib.ue_gen_comm_init(world, rank, bytes(t.cpu().numpy().tobytes())) != 0:
        raise RuntimeError(lib.ue_gen_last_error().decode())
    torch.cuda.synchronize()
    dist.barrier()
