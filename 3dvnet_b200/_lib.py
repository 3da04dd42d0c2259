"""ctypes binding of lib3dvnet_b200.so — the C ABI declared in include/dv3d.h.

The prototypes are parsed from the header itself, so the binding cannot drift from it.
There is NO fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'dv3d.h')
LIB_PATH = os.path.join(HERE, 'lib3dvnet_b200.so')

DV3D_OK, DV3D_EINVAL, DV3D_ECUDA, DV3D_ENOSPC = 0, -1, -2, -3


class Dv3dError(RuntimeError):
    def __init__(self, code, fn, msg):
        super().__init__('%s failed (%d): %s' % (fn, code, msg))
        self.code = code


class VoxelGrid(ctypes.Structure):
    """dv3d_voxel_grid_t"""
    _fields_ = [('bbox_min', ctypes.c_float * 3), ('bbox_max', ctypes.c_float * 3), ('edge_len', ctypes.c_float),
                ('n_cells', ctypes.c_longlong * 3), ('grid_size', ctypes.c_longlong * 3),
                ('n_batch', ctypes.c_longlong), ('total_cells', ctypes.c_longlong)]


_SCALARS = {'int': ctypes.c_int, 'long long': ctypes.c_longlong, 'float': ctypes.c_float, 'double': ctypes.c_double,
            'size_t': ctypes.c_size_t}
_RET = dict(_SCALARS, **{'const char*': ctypes.c_char_p})


def parse_header(path=HEADER):
    """-> {name: (restype, [(ctype, argname), ...])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'//[^\n]*', '', src)
    src = re.sub(r'typedef\s+struct.*?\}\s*\w+\s*;', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'(const char\*|size_t|long long|int)\s+(dv3d_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                am = re.match(r'^(.*?)(\w+)$', a)
                typ, an = am.group(1).strip(), am.group(2)
                if '*' in typ:
                    parsed.append((ctypes.c_void_p, an))
                else:
                    parsed.append((_SCALARS[typ.replace('const ', '')], an))
        protos[name] = (_RET[ret], parsed)
    return protos


class _Lib(object):
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError('%s is missing: build it with `python 3dvnet_b200/build.py` (no CPU fallback exists)'
                              % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (ret, args) in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = ret
            fn.argtypes = [t for t, _ in args]
        assert self.cdll.dv3d_abi_version() == 4

    def last_error(self):
        return self.cdll.dv3d_last_error().decode()

    def call(self, name, *args):
        """Call an int-returning entry point and raise Dv3dError on a negative code."""
        rc = getattr(self.cdll, name)(*args)
        if rc != DV3D_OK:
            raise Dv3dError(rc, name, self.last_error())

    def raw(self, name):
        return getattr(self.cdll, name)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
