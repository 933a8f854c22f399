"""Host-side vector/matrix value types of the zdr Python surface.

The reference re-exports LuisaCompute's ``float3`` (``__init__.py:1``, ``render.py:10``) and
users build instance transforms with ``luisa.float4x4`` (``test_lightstage.py:44``, column-major
16-scalar constructor; ``render.py:84`` uses ``float4x4(1.0)`` for identity).  These are the
same value types without LuisaCompute: plain Python, no device state.
"""
from __future__ import annotations

import numbers
from typing import Iterable, Sequence


def _is_num(x) -> bool:
    return isinstance(x, numbers.Real)


class _Vec:
    __slots__ = ("_v",)
    _n = 0
    _names = "xyzw"

    def __init__(self, *args):
        n = self._n
        if len(args) == 0:
            vals = [self._cast(0)] * n
        elif len(args) == 1 and _is_num(args[0]):
            vals = [self._cast(args[0])] * n                       # float3(0.0) broadcast
        elif len(args) == 1:
            vals = [self._cast(c) for c in args[0]]
        else:
            vals = []
            for a in args:                                           # float4(float3, 1.0) style
                if _is_num(a):
                    vals.append(self._cast(a))
                else:
                    vals.extend(self._cast(c) for c in a)
        if len(vals) != n:
            raise TypeError(f"{type(self).__name__} needs {n} components, got {len(vals)}")
        object.__setattr__(self, "_v", vals)

    @staticmethod
    def _cast(x):
        return float(x)

    def __iter__(self):
        return iter(self._v)

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return self._v[i]

    def __setitem__(self, i, val):
        self._v[i] = self._cast(val)

    def __getattr__(self, name):
        idx = self._names.find(name) if len(name) == 1 else -1
        if 0 <= idx < self._n:
            return self._v[idx]
        raise AttributeError(name)

    def __setattr__(self, name, val):
        idx = self._names.find(name) if len(name) == 1 else -1
        if 0 <= idx < self._n:
            self._v[idx] = self._cast(val)
        else:
            object.__setattr__(self, name, val)

    def __eq__(self, other):
        try:
            return type(other) is type(self) and list(other) == self._v
        except TypeError:
            return False

    def __hash__(self):
        return hash((type(self).__name__, tuple(self._v)))

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(repr(c) for c in self._v)})"

    def _binary(self, other, op):
        if _is_num(other):
            return type(self)(*[op(a, other) for a in self._v])
        return type(self)(*[op(a, b) for a, b in zip(self._v, other)])

    def __add__(self, o):
        return self._binary(o, lambda a, b: a + b)

    def __sub__(self, o):
        return self._binary(o, lambda a, b: a - b)

    def __mul__(self, o):
        return self._binary(o, lambda a, b: a * b)

    __rmul__ = __mul__
    __radd__ = __add__

    def __truediv__(self, o):
        return self._binary(o, lambda a, b: a / b)

    def __neg__(self):
        return type(self)(*[-a for a in self._v])

    def copy(self):
        return type(self)(*self._v)


class float2(_Vec):
    _n = 2


class float3(_Vec):
    _n = 3


class float4(_Vec):
    _n = 4


class int2(_Vec):
    _n = 2

    @staticmethod
    def _cast(x):
        return int(x)


class float4x4:
    """Column-major 4x4 matrix: ``float4x4(*m.transpose().flatten())`` as at
    ``test_lightstage.py:44``; ``float4x4(1.0)`` is the identity (``render.py:84``)."""

    __slots__ = ("cols",)

    def __init__(self, *args):
        if len(args) == 1 and _is_num(args[0]):
            s = float(args[0])
            self.cols = [float4(s if r == c else 0.0 for r in range(4)) for c in range(4)]
        elif len(args) == 16:
            self.cols = [float4(*args[4 * c:4 * c + 4]) for c in range(4)]
        elif len(args) == 4:
            self.cols = [float4(a) for a in args]
        elif len(args) == 1:
            flat = [float(x) for row in args[0] for x in (row if not _is_num(row) else [row])]
            if len(flat) != 16:
                raise TypeError("float4x4 needs 16 scalars")
            self.cols = [float4(*flat[4 * c:4 * c + 4]) for c in range(4)]
        else:
            raise TypeError("float4x4(scalar) | float4x4(16 scalars, column-major) | float4x4(4 columns)")

    def __getitem__(self, c):
        return self.cols[c]

    def flat(self) -> list:
        """16 floats, column-major (the layout the C ABI takes)."""
        return [x for col in self.cols for x in col]

    @classmethod
    def from_rows(cls, rows: Sequence[Iterable[float]]):
        r = [list(map(float, row)) for row in rows]
        return cls(*[r[i][c] for c in range(4) for i in range(4)])

    def __repr__(self):
        return f"float4x4({', '.join(repr(x) for x in self.flat())})"


class Camera:
    """``Camera = luisa.StructType(fov=float, origin=float3, target=float3, up=float3)``
    (``render.py:28``). ``fov`` is the full horizontal angle in radians (``camera.py:15``)."""

    __slots__ = ("fov", "origin", "target", "up")

    def __init__(self, fov=40 / 180 * 3.1415926, origin=None, target=None, up=None):
        self.fov = float(fov)
        self.origin = float3(origin) if origin is not None else float3(1.0, 0.5, 0.0)
        self.target = float3(target) if target is not None else float3(0.0, 0.0, 0.0)
        self.up = float3(up) if up is not None else float3(0.0, 1.0, 0.0)

    def copy(self):
        return Camera(self.fov, self.origin.copy(), self.target.copy(), self.up.copy())

    def __repr__(self):
        return f"Camera(fov={self.fov!r}, origin={self.origin!r}, target={self.target!r}, up={self.up!r})"
