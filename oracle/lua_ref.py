"""TEST INFRASTRUCTURE ONLY -- the reference's own Lua interpreter (aotus/external/lua-5.4.8,
compiled from where it lies by `make -C oracle ref` into oracle/_ref/liblua_ref.so) behind ctypes,
to evaluate the reference's configuration scripts (musubi.lua) the way aotus does: run the chunk,
read globals / nested table entries, call Lua functions (space-time functions of initial and
boundary conditions).  Used by tests/ to check that the restated golden cases hold the values the
reference's own scripts define.  Nothing under musubi_b200/ may import this module."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "liblua_ref.so")

LUA_OK, LUA_MULTRET = 0, -1
LUA_TNIL, LUA_TBOOLEAN, LUA_TNUMBER, LUA_TSTRING, LUA_TTABLE, LUA_TFUNCTION = 0, 1, 3, 4, 5, 6


def available():
    return os.path.exists(_PATH)


def _load():
    L = ctypes.CDLL(_PATH)
    vp, ci, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p
    sig = {
        "luaL_newstate": (vp, []), "luaL_openlibs": (None, [vp]), "lua_close": (None, [vp]),
        "luaL_loadstring": (ci, [vp, cs]), "luaL_loadfilex": (ci, [vp, cs, cs]),
        "lua_pcallk": (ci, [vp, ci, ci, ci, ctypes.c_ssize_t, vp]),
        "lua_getglobal": (ci, [vp, cs]), "lua_getfield": (ci, [vp, ci, cs]),
        "lua_geti": (ci, [vp, ci, ctypes.c_longlong]), "lua_settop": (None, [vp, ci]), "lua_gettop": (ci, [vp]),
        "lua_type": (ci, [vp, ci]), "lua_tonumberx": (ctypes.c_double, [vp, ci, ctypes.POINTER(ci)]),
        "lua_tointegerx": (ctypes.c_longlong, [vp, ci, ctypes.POINTER(ci)]), "lua_isinteger": (ci, [vp, ci]),
        "lua_toboolean": (ci, [vp, ci]), "lua_tolstring": (cs, [vp, ci, ctypes.POINTER(ctypes.c_size_t)]),
        "lua_pushnumber": (None, [vp, ctypes.c_double]), "lua_pushnil": (None, [vp]),
        "lua_next": (ci, [vp, ci]), "lua_rawlen": (ctypes.c_size_t, [vp, ci]), "lua_pushvalue": (None, [vp, ci]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return L


class LuaScript:
    """a Lua state holding the globals of one executed script"""

    def __init__(self, path=None, text=None, chdir=True):
        if not available():
            raise RuntimeError("oracle/_ref/liblua_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = _load()
        self.L = self.lib.luaL_newstate()
        self.lib.luaL_openlibs(self.L)
        cwd = os.getcwd()
        try:
            if path is not None:
                if chdir:                       # scripts `require` their neighbours
                    os.chdir(os.path.dirname(os.path.abspath(path)))
                rc = self.lib.luaL_loadfilex(self.L, os.path.abspath(path).encode(), None)
            else:
                rc = self.lib.luaL_loadstring(self.L, text.encode())
            if rc == LUA_OK:
                rc = self.lib.lua_pcallk(self.L, 0, 0, 0, 0, None)
        finally:
            os.chdir(cwd)
        if rc != LUA_OK:
            msg = self.lib.lua_tolstring(self.L, -1, None)
            raise RuntimeError("Lua: %s" % (msg.decode() if msg else rc))

    def close(self):
        if self.L:
            self.lib.lua_close(self.L)
            self.L = None

    # -- stack helpers ----------------------------------------------------------------
    def _push_path(self, path):
        """push global `a.b[2].c` given as 'a.b.2.c'; returns stack top before the push"""
        top = self.lib.lua_gettop(self.L)
        parts = path.split(".")
        self.lib.lua_getglobal(self.L, parts[0].encode())
        for p in parts[1:]:
            if self.lib.lua_type(self.L, -1) != LUA_TTABLE:
                self.lib.lua_settop(self.L, top)
                self.lib.lua_pushnil(self.L)
                break
            if p.isdigit():
                self.lib.lua_geti(self.L, -1, int(p))
            else:
                self.lib.lua_getfield(self.L, -1, p.encode())
        return top

    def _value(self, idx=-1):
        t = self.lib.lua_type(self.L, idx)
        if t == LUA_TNIL:
            return None
        if t == LUA_TBOOLEAN:
            return bool(self.lib.lua_toboolean(self.L, idx))
        if t == LUA_TNUMBER:
            if self.lib.lua_isinteger(self.L, idx):
                return int(self.lib.lua_tointegerx(self.L, idx, None))
            return float(self.lib.lua_tonumberx(self.L, idx, None))
        if t == LUA_TSTRING:
            return self.lib.lua_tolstring(self.L, idx, None).decode()
        if t == LUA_TFUNCTION:
            return "<function>"
        if t == LUA_TTABLE:
            self.lib.lua_pushvalue(self.L, idx)
            out = {}
            self.lib.lua_pushnil(self.L)
            while self.lib.lua_next(self.L, -2):
                kt = self.lib.lua_type(self.L, -2)
                if kt == LUA_TNUMBER:
                    key = int(self.lib.lua_tonumberx(self.L, -2, None))
                else:
                    self.lib.lua_pushvalue(self.L, -2)      # tolstring on a copy: never convert the key in place
                    key = self.lib.lua_tolstring(self.L, -1, None).decode()
                    self.lib.lua_settop(self.L, -2)
                out[key] = self._value(-1)
                self.lib.lua_settop(self.L, -2)
            self.lib.lua_settop(self.L, -2)
            n = len(out)
            if n and all(isinstance(k, int) for k in out) and sorted(out) == list(range(1, n + 1)):
                return [out[k] for k in range(1, n + 1)]
            return out
        return "<%d>" % t

    def get(self, path):
        """the value at `a.b.2.c` (tables -> dict, or list when the keys are 1..n)"""
        top = self._push_path(path)
        v = self._value(-1)
        self.lib.lua_settop(self.L, top)
        return v

    def call(self, path, *args):
        """call the Lua function at `path` with numbers; one result: number, or table -> list"""
        top = self._push_path(path)
        if self.lib.lua_type(self.L, -1) != LUA_TFUNCTION:
            self.lib.lua_settop(self.L, top)
            raise TypeError("%s is not a Lua function" % path)
        # the function must be the only thing above `top`
        self.lib.lua_pushvalue(self.L, -1)
        for a in args:
            self.lib.lua_pushnumber(self.L, float(a))
        rc = self.lib.lua_pcallk(self.L, len(args), 1, 0, 0, None)
        if rc != LUA_OK:
            msg = self.lib.lua_tolstring(self.L, -1, None)
            self.lib.lua_settop(self.L, top)
            raise RuntimeError("Lua: %s" % (msg.decode() if msg else rc))
        v = self._value(-1)
        self.lib.lua_settop(self.L, top)
        return v
