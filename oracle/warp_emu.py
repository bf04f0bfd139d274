"""warp_emu -- a sequential pure-Python / numpy stand-in for the part of NVIDIA Warp's API that
/root/reference/warp_mpm uses.

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/): it exists so that the reference's OWN
kernel source (warp_mpm/mpm_utils.py, mpm_solver.py, mpm_data_structure.py, unmodified, read from
/root/reference at fixture-generation time) can be executed in the build container, where warp-lang
0.10.1 is not installable, to produce the golden vectors under tests/golden/ (tests/golden/make_golden.py).
Nothing in the product path imports it.

What is emulated faithfully (plain semantics of the Warp language):
  * @wp.kernel / @wp.func are ordinary Python functions; wp.launch runs the kernel once per thread index,
    in increasing index order, with wp.tid() returning the index (sequential == one legal schedule of the
    reference's atomics);
  * wp.vec3 / wp.mat33 value types: `*` is matrix-matrix / matrix-vector / scalar product as in Warp,
    wp.mat33(v0, v1, v2) builds the matrix from COLUMN vectors, wp.mat33(9 scalars) is row-major;
  * wp.array element reads return copies (value semantics), wp.atomic_add adds in place;
  * wp.int truncates toward zero, wp.normalize(0) = 0.
What is ASSUMED about the two third-party numerical routines whose source is not available offline
(SURVEY.md section 8c): wp.qr3 returns a rotation Q (det +1, built here from three Givens rotations) and
upper-triangular R whose diagonal signs are arbitrary -- the reference's own sign fix-ups
(mpm_utils.py:112-123, 184-195) are therefore exercised; wp.svd3 returns U, V with det +1 and
sigma_0 >= sigma_1 >= |sigma_2| (sign on the smallest singular value), the convention of the McAdams
solver that warp-lang wraps.
Arithmetic runs in the dtype chosen with set_precision ("f64" for algorithm pinning, "f32" mimics Warp's
fp32 storage and arithmetic through numpy's weak-scalar promotion).
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np

_DT = np.float64
_TID = 0
_MESHES = {}


def set_precision(p: str) -> None:
    global _DT
    _DT = np.float64 if p == "f64" else np.float32


def real():
    return _DT


# ------------------------------------------------------------------ value types
class _Val(np.ndarray):
    def __array_finalize__(self, obj):
        pass

    def __reduce__(self):
        return (np.asarray, (np.asarray(self),))


class vec3(_Val):
    _shape_ = (3,)

    def __new__(cls, *a):
        o = np.zeros(3, dtype=_DT).view(cls)
        if len(a) == 1:
            o[:] = a[0]
        elif len(a) == 3:
            o[0], o[1], o[2] = a
        elif len(a) != 0:
            raise TypeError("vec3 takes 0, 1 or 3 arguments")
        return o

    def __mul__(self, other):
        if isinstance(other, (vec3, mat33)):
            raise TypeError("use wp.cw_mul / wp.dot for vector products")
        return np.multiply(self, other).view(vec3)

    __rmul__ = __mul__


class vec2(_Val):
    _shape_ = (2,)

    def __new__(cls, *a):
        o = np.zeros(2, dtype=_DT).view(cls)
        if len(a) == 1:
            o[:] = a[0]
        elif len(a) == 2:
            o[0], o[1] = a
        return o


class mat33(_Val):
    _shape_ = (3, 3)

    def __new__(cls, *a):
        o = np.zeros((3, 3), dtype=_DT).view(cls)
        if len(a) == 1:
            o[:] = a[0]
        elif len(a) == 3:  # column vectors
            for c in range(3):
                o[:, c] = np.asarray(a[c])
        elif len(a) == 9:
            o[:] = np.asarray(a, dtype=_DT).reshape(3, 3)
        elif len(a) != 0:
            raise TypeError("mat33 takes 0, 1, 3 or 9 arguments")
        return o

    def __mul__(self, other):
        if isinstance(other, mat33):
            return np.matmul(np.asarray(self), np.asarray(other)).view(mat33)
        if isinstance(other, vec3):
            return np.matmul(np.asarray(self), np.asarray(other)).view(vec3)
        return np.multiply(np.asarray(self), other).view(mat33)

    def __rmul__(self, other):
        if isinstance(other, vec3):
            return np.matmul(np.asarray(other), np.asarray(self)).view(vec3)
        return np.multiply(np.asarray(self), other).view(mat33)

    def __matmul__(self, other):
        return self.__mul__(other)


class mat22(_Val):
    _shape_ = (2, 2)

    def __new__(cls, *a):
        o = np.zeros((2, 2), dtype=_DT).view(cls)
        if len(a) == 1:
            o[:] = a[0]
        elif len(a) == 4:
            o[:] = np.asarray(a, dtype=_DT).reshape(2, 2)
        return o

    def __mul__(self, other):
        if isinstance(other, mat22):
            return np.matmul(np.asarray(self), np.asarray(other)).view(mat22)
        if isinstance(other, vec2):
            return np.matmul(np.asarray(self), np.asarray(other)).view(vec2)
        return np.multiply(np.asarray(self), other).view(mat22)

    __rmul__ = __mul__


float32 = float
float64 = float
int32 = int
uint64 = int


def float_(x):  # wp.float(x)
    return _DT(x)


def int_(x):  # wp.int(x): truncation toward zero
    return int(x)


# ------------------------------------------------------------------ arrays
def _inner(dtype):
    return getattr(dtype, "_shape_", ())


class array:
    """wp.array; also callable as an annotation placeholder: wp.array(dtype=..., ndim=...)."""

    def __init__(self, data=None, dtype=float, ndim=None, **kw):
        self.dtype = dtype
        self.np = None
        if data is not None:
            base = int if dtype in (int, np.int32) else _DT
            self.np = np.array(data, dtype=base)
        self.requires_grad = kw.get("requires_grad", False)

    @property
    def shape(self):
        k = len(_inner(self.dtype))
        return self.np.shape[: self.np.ndim - k] if k else self.np.shape

    def numpy(self):
        return self.np

    def _wrap(self, v):
        if self.dtype in (vec3, mat33, vec2, mat22):
            return np.array(v, dtype=_DT).view(self.dtype)
        return v

    def __getitem__(self, idx):
        return self._wrap(self.np[idx])

    def __setitem__(self, idx, val):
        self.np[idx] = np.asarray(val)

    def __len__(self):
        return self.shape[0]


def _alloc(shape, dtype):
    shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
    a = array(dtype=dtype)
    base = int if dtype in (int, np.int32) else _DT
    a.np = np.zeros(shape + tuple(_inner(dtype)), dtype=base)
    return a


def zeros(shape=None, dtype=float, device=None, requires_grad=False, **kw):
    return _alloc(shape, dtype)


empty = zeros


def zeros_like(a, requires_grad=False, **kw):
    b = array(dtype=a.dtype)
    b.np = np.zeros_like(a.np)
    return b


def from_numpy(arr, dtype=float, device=None, requires_grad=False, **kw):
    a = array(dtype=dtype)
    base = int if dtype in (int, np.int32) else _DT
    a.np = np.array(arr, dtype=base)
    k = _inner(dtype)
    if k and tuple(a.np.shape[-len(k):]) != tuple(k):
        a.np = a.np.reshape((-1,) + tuple(k))
    return a


def from_torch(t, dtype=float, requires_grad=None, **kw):
    return from_numpy(t.detach().cpu().numpy(), dtype=dtype)


def to_torch(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a.np))


def copy(dest, src, **kw):
    dest.np[...] = src.np


def atomic_add(arr, *a):
    *idx, val = a
    idx = tuple(idx) if len(idx) > 1 else idx[0]
    old = arr[idx]
    arr.np[idx] = arr.np[idx] + np.asarray(val)
    return old


def atomic_sub(arr, *a):
    *idx, val = a
    idx = tuple(idx) if len(idx) > 1 else idx[0]
    old = arr[idx]
    arr.np[idx] = arr.np[idx] - np.asarray(val)
    return old


# ------------------------------------------------------------------ decorators / launch
def kernel(f=None, **kw):
    return f if f is not None else (lambda g: g)


func = kernel


def struct(cls):
    """@wp.struct: annotated fields exist on every instance, zero-initialised as in Warp."""
    ann = dict(getattr(cls, "__annotations__", {}))
    orig = cls.__init__

    def __init__(self, *a, **kw):
        for name, t in ann.items():
            if isinstance(t, array):
                setattr(self, name, None)
            elif t in (vec3, mat33, vec2, mat22):
                setattr(self, name, t())
            elif t is int or t is int32:
                setattr(self, name, 0)
            else:
                setattr(self, name, 0.0)
        orig(self, *a, **kw)
    cls.__init__ = __init__
    return cls


def tid():
    return _TID


def launch(kernel=None, dim=None, inputs=(), outputs=(), device=None, **kw):
    global _TID
    args = list(inputs) + list(outputs)
    if isinstance(dim, (int, np.integer)):
        for i in range(int(dim)):
            _TID = i
            kernel(*args)
    else:
        dims = [int(d) for d in dim]
        for idx in np.ndindex(*dims):
            _TID = idx if len(dims) > 1 else idx[0]
            kernel(*args)


class ScopedTimer:
    def __init__(self, name="", *a, **kw):
        self.dict = kw.get("dict")
        self.name = name

    def __enter__(self):
        return self

    def __exit__(self, *a):
        if self.dict is not None:
            self.dict.setdefault(self.name, []).append(0.0)
        return False


def init():
    pass


def synchronize():
    pass


class _Cfg:
    mode = "release"
    verify_cuda = False


config = _Cfg()
context = types.SimpleNamespace(Devicelike=object, runtime=None)
types_ = types.SimpleNamespace(array=array, float32=float, int32=int)


# ------------------------------------------------------------------ builtins
def transpose(m):
    return np.array(np.asarray(m).T, dtype=_DT).view(type(m))


def determinant(m):
    a = np.asarray(m)
    if a.shape == (2, 2):
        return a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]
    return (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0])
            + a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))


def inverse(m):
    return np.linalg.inv(np.asarray(m, dtype=np.float64)).astype(_DT).view(type(m))


def outer(a, b):
    return np.outer(np.asarray(a), np.asarray(b)).astype(_DT).view(mat33 if len(a) == 3 else mat22)


def cw_mul(a, b):
    return np.multiply(np.asarray(a), np.asarray(b)).view(type(a))


def cw_div(a, b):
    return np.divide(np.asarray(a), np.asarray(b)).view(type(a))


def ddot(a, b):
    return (np.asarray(a) * np.asarray(b)).sum(dtype=_DT)


def diag(v):
    return np.diag(np.asarray(v)).astype(_DT).view(mat33)


def dot(a, b):
    a, b = np.asarray(a), np.asarray(b)
    s = a[0] * b[0]
    for i in range(1, len(a)):
        s = s + a[i] * b[i]
    return s


def cross(a, b):
    return vec3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def length(a):
    return np.sqrt(dot(a, a))


def normalize(a):
    n = length(a)
    if n > 0.0:
        return (np.asarray(a) / n).view(type(a))
    return np.zeros_like(np.asarray(a)).view(type(a))


def clamp(x, a, b):
    return min(max(x, a), b)


def min_(a, b):
    return a if a < b else b


def max_(a, b):
    return a if a > b else b


def _u(f):
    def g(x):
        return _DT(f(x))
    return g


sqrt = lambda x: np.sqrt(_DT(x))
exp = lambda x: np.exp(_DT(x))
log = lambda x: np.log(_DT(x)) if x > 0 else _DT(-np.inf)
sin = lambda x: np.sin(_DT(x))
cos = lambda x: np.cos(_DT(x))
tan = lambda x: np.tan(_DT(x))
acos = lambda x: np.arccos(_DT(x))
pow_ = lambda x, y: np.power(_DT(x), _DT(y))
abs_ = lambda x: abs(x)


def _givens(a, b):
    r = math.hypot(a, b)
    if r == 0.0:
        return 1.0, 0.0
    return a / r, b / r


def qr3(A, Q, R):
    """Givens QR: Q rotation (det +1), R upper triangular; diagonal signs left as they fall."""
    Rm = np.array(A, dtype=np.float64)
    Qm = np.eye(3)
    for (i, j) in ((1, 0), (2, 0), (2, 1)):  # zero R[i,j] with a rotation in the (j,i) plane
        c, s = _givens(Rm[j, j], Rm[i, j])
        G = np.eye(3)
        G[j, j], G[j, i], G[i, j], G[i, i] = c, s, -s, c
        Rm = G @ Rm
        Qm = Qm @ G.T
    Rm[1, 0] = Rm[2, 0] = Rm[2, 1] = 0.0
    Q[:] = Qm.astype(_DT)
    R[:] = Rm.astype(_DT)


def svd3(A, U, S, V):
    """A = U diag(S) V^T with det U = det V = +1, S0 >= S1 >= |S2| (sign on the smallest)."""
    u, s, vt = np.linalg.svd(np.array(A, dtype=np.float64))
    v = vt.T
    if np.linalg.det(u) < 0:
        u[:, 2] = -u[:, 2]
        s[2] = -s[2]
    if np.linalg.det(v) < 0:
        v[:, 2] = -v[:, 2]
        s[2] = -s[2]
    U[:] = u.astype(_DT)
    S[:] = s.astype(_DT)
    V[:] = v.astype(_DT)


# ------------------------------------------------------------------ meshes
class Mesh:
    def __init__(self, points=None, velocities=None, indices=None, **kw):
        self.points, self.velocities, self.indices = points, velocities, indices
        self.id = len(_MESHES) + 1
        _MESHES[self.id] = self

    def refit(self):
        pass


def mesh_get(mesh_id):
    return _MESHES[mesh_id]


def mesh_eval_face_normal(mesh_id, face):
    m = _MESHES[mesh_id]
    i0, i1, i2 = (int(m.indices[3 * face + k]) for k in range(3))
    p, q, r = m.points[i0], m.points[i1], m.points[i2]
    return normalize(cross(q - p, r - p))


# ------------------------------------------------------------------ module installation
def install():
    """Register this module as `warp` (+ `warp.torch`), and shims for the two imports of the reference that
    are pure plumbing (`warp_utils.from_torch_safe`, `jaxtyping`)."""
    me = sys.modules[__name__]
    ns = types.ModuleType("warp")
    for k, v in vars(me).items():
        if not k.startswith("__"):
            setattr(ns, k, v)
    # Warp names that collide with Python builtins / keywords in this file
    ns.float = float_
    ns.int = int_
    ns.min = min_
    ns.max = max_
    ns.abs = abs_
    ns.pow = pow_
    ns.types = types_
    wt = types.ModuleType("warp.torch")
    wt.from_torch = from_torch
    wt.to_torch = to_torch
    ns.torch = wt
    sys.modules["warp"] = ns
    sys.modules["warp.torch"] = wt
    wu = types.ModuleType("warp_utils")

    def from_torch_safe(t, dtype=None, requires_grad=None, grad=None):
        a = from_torch(t, dtype=dtype if dtype is not None else float)
        a._tensor = t
        return a
    wu.from_torch_safe = from_torch_safe
    sys.modules["warp_utils"] = wu
    jt = types.ModuleType("jaxtyping")

    class _Sub:  # Float[Tensor, "n"] | Float[Tensor, "1"] in annotations
        def __class_getitem__(cls, item):
            return cls

        def __or__(self, other):
            return self
    jt.Float = jt.Int = jt.Shaped = _Sub
    sys.modules["jaxtyping"] = jt
    return ns
