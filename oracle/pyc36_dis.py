"""Disassembler for CPython 3.6 ``.pyc`` files that runs on any later Python.  TEST INFRASTRUCTURE ONLY.

The reference's genetic algorithm survives only as
``code/__pycache__/updated_deap_class.cpython-36.pyc``; Python 3.12's ``marshal`` cannot build a 3.6
code object, so this module reads the marshal stream itself (version 4 with FLAG_REF) and prints the
3.6 wordcode (two bytes per instruction, ``EXTENDED_ARG`` prefixes) with constants, names and line
numbers resolved.  ``python -m oracle.pyc36_dis <file.pyc> > listing.txt``; the listings the GA
restatement in ``scbonita2_b200/genetic_algorithm.py`` was written from are committed under
``oracle/bytecode/``.
"""
from __future__ import annotations

import struct
import sys

OPNAMES = {
    1: "POP_TOP", 2: "ROT_TWO", 3: "ROT_THREE", 4: "DUP_TOP", 5: "DUP_TOP_TWO", 9: "NOP", 10: "UNARY_POSITIVE",
    11: "UNARY_NEGATIVE", 12: "UNARY_NOT", 15: "UNARY_INVERT", 16: "BINARY_MATRIX_MULTIPLY",
    17: "INPLACE_MATRIX_MULTIPLY", 19: "BINARY_POWER", 20: "BINARY_MULTIPLY", 22: "BINARY_MODULO", 23: "BINARY_ADD",
    24: "BINARY_SUBTRACT", 25: "BINARY_SUBSCR", 26: "BINARY_FLOOR_DIVIDE", 27: "BINARY_TRUE_DIVIDE",
    28: "INPLACE_FLOOR_DIVIDE", 29: "INPLACE_TRUE_DIVIDE", 50: "GET_AITER", 51: "GET_ANEXT", 52: "BEFORE_ASYNC_WITH",
    55: "INPLACE_ADD", 56: "INPLACE_SUBTRACT", 57: "INPLACE_MULTIPLY", 59: "INPLACE_MODULO", 60: "STORE_SUBSCR",
    61: "DELETE_SUBSCR", 62: "BINARY_LSHIFT", 63: "BINARY_RSHIFT", 64: "BINARY_AND", 65: "BINARY_XOR", 66: "BINARY_OR",
    67: "INPLACE_POWER", 68: "GET_ITER", 69: "GET_YIELD_FROM_ITER", 70: "PRINT_EXPR", 71: "LOAD_BUILD_CLASS",
    72: "YIELD_FROM", 73: "GET_AWAITABLE", 75: "INPLACE_LSHIFT", 76: "INPLACE_RSHIFT", 77: "INPLACE_AND",
    78: "INPLACE_XOR", 79: "INPLACE_OR", 80: "BREAK_LOOP", 81: "WITH_CLEANUP_START", 82: "WITH_CLEANUP_FINISH",
    83: "RETURN_VALUE", 84: "IMPORT_STAR", 85: "SETUP_ANNOTATIONS", 86: "YIELD_VALUE", 87: "POP_BLOCK",
    88: "END_FINALLY", 89: "POP_EXCEPT", 90: "STORE_NAME", 91: "DELETE_NAME", 92: "UNPACK_SEQUENCE", 93: "FOR_ITER",
    94: "UNPACK_EX", 95: "STORE_ATTR", 96: "DELETE_ATTR", 97: "STORE_GLOBAL", 98: "DELETE_GLOBAL", 100: "LOAD_CONST",
    101: "LOAD_NAME", 102: "BUILD_TUPLE", 103: "BUILD_LIST", 104: "BUILD_SET", 105: "BUILD_MAP", 106: "LOAD_ATTR",
    107: "COMPARE_OP", 108: "IMPORT_NAME", 109: "IMPORT_FROM", 110: "JUMP_FORWARD", 111: "JUMP_IF_FALSE_OR_POP",
    112: "JUMP_IF_TRUE_OR_POP", 113: "JUMP_ABSOLUTE", 114: "POP_JUMP_IF_FALSE", 115: "POP_JUMP_IF_TRUE",
    116: "LOAD_GLOBAL", 119: "CONTINUE_LOOP", 120: "SETUP_LOOP", 121: "SETUP_EXCEPT", 122: "SETUP_FINALLY",
    124: "LOAD_FAST", 125: "STORE_FAST", 126: "DELETE_FAST", 127: "STORE_ANNOTATION", 130: "RAISE_VARARGS",
    131: "CALL_FUNCTION", 132: "MAKE_FUNCTION", 133: "BUILD_SLICE", 135: "LOAD_CLOSURE", 136: "LOAD_DEREF",
    137: "STORE_DEREF", 138: "DELETE_DEREF", 141: "CALL_FUNCTION_KW", 142: "CALL_FUNCTION_EX", 143: "SETUP_WITH",
    144: "EXTENDED_ARG", 145: "LIST_APPEND", 146: "SET_ADD", 147: "MAP_ADD", 148: "LOAD_CLASSDEREF",
    149: "BUILD_LIST_UNPACK", 150: "BUILD_MAP_UNPACK", 151: "BUILD_MAP_UNPACK_WITH_CALL", 152: "BUILD_TUPLE_UNPACK",
    153: "BUILD_SET_UNPACK", 154: "SETUP_ASYNC_WITH", 155: "FORMAT_VALUE", 156: "BUILD_CONST_KEY_MAP",
    157: "BUILD_STRING", 158: "BUILD_TUPLE_UNPACK_WITH_CALL"}
HAVE_ARGUMENT = 90
HASCONST, HASNAME = {100}, {90, 91, 95, 96, 97, 98, 101, 106, 108, 109, 116}
HASLOCAL, HASFREE = {124, 125, 126}, {135, 136, 137, 138, 148}
HASJREL, HASJABS = {93, 110, 120, 121, 122, 143, 154}, {111, 112, 113, 114, 115, 119}
CMP_OP = ('<', '<=', '==', '!=', '>', '>=', 'in', 'not in', 'is', 'is not', 'exception match', 'BAD')


class Code36:
    """A 3.6 code object as plain data."""
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __repr__(self):
        return f"<code object {self.name} at line {self.firstlineno}>"


class Reader:
    def __init__(self, data: bytes):
        self.data, self.pos, self.refs = data, 0, []

    def _take(self, n):
        chunk = self.data[self.pos:self.pos + n]
        self.pos += n
        return chunk

    def _i32(self):
        return struct.unpack("<i", self._take(4))[0]

    def load(self):
        code = self._take(1)[0]
        flag, kind = code & 0x80, chr(code & 0x7F)
        slot = None
        if flag and kind in "(c[{<>)":          # containers register before their contents are read
            slot = len(self.refs)
            self.refs.append(None)
        value = self._load(kind)
        if slot is not None:
            self.refs[slot] = value
        elif flag:
            self.refs.append(value)
        return value

    def _load(self, kind):
        if kind == '0':
            return None
        if kind == 'N':
            return None
        if kind == 'F':
            return False
        if kind == 'T':
            return True
        if kind == 'S':
            return StopIteration
        if kind == '.':
            return Ellipsis
        if kind == 'i':
            return self._i32()
        if kind == 'l':
            n = self._i32()
            digits = struct.unpack(f"<{abs(n)}H", self._take(2 * abs(n)))
            value = sum(d << (15 * k) for k, d in enumerate(digits))
            return -value if n < 0 else value
        if kind == 'g':
            return struct.unpack("<d", self._take(8))[0]
        if kind == 'f':
            return float(self._take(self._take(1)[0]).decode())
        if kind == 'y':
            return complex(*struct.unpack("<dd", self._take(16)))
        if kind == 's':
            return bytes(self._take(self._i32()))
        if kind in 'tu':
            return self._take(self._i32()).decode("utf-8", "surrogatepass")
        if kind in 'aA':
            return self._take(self._i32()).decode("latin-1")
        if kind in 'zZ':
            return self._take(self._take(1)[0]).decode("latin-1")
        if kind == ')':
            return tuple(self.load() for _ in range(self._take(1)[0]))
        if kind == '(':
            return tuple(self.load() for _ in range(self._i32()))
        if kind == '[':
            return [self.load() for _ in range(self._i32())]
        if kind in '<>':
            return frozenset(self.load() for _ in range(self._i32()))
        if kind == '{':
            out = {}
            while True:
                key = self.load()
                if key is None and self.data[self.pos - 1] == ord('0'):
                    break
                out[key] = self.load()
            return out
        if kind == 'r':
            return self.refs[self._i32()]
        if kind == 'c':
            argcount, kwonly, nlocals, stacksize, flags = (self._i32() for _ in range(5))
            code = self.load()
            consts, names, varnames, freevars, cellvars = (self.load() for _ in range(5))
            filename, name = self.load(), self.load()
            firstlineno = self._i32()
            lnotab = self.load()
            return Code36(argcount=argcount, kwonlyargcount=kwonly, nlocals=nlocals, stacksize=stacksize,
                          flags=flags, code=code, consts=consts, names=names, varnames=varnames,
                          freevars=freevars, cellvars=cellvars, filename=filename, name=name,
                          firstlineno=firstlineno, lnotab=lnotab)
        raise ValueError(f"unknown marshal type {kind!r} at byte {self.pos - 1}")


def line_starts(co: Code36) -> dict:
    """{bytecode offset: source line} from the 3.6 lnotab (signed line increments)."""
    out, addr, line = {0: co.firstlineno}, 0, co.firstlineno
    tab = co.lnotab
    for k in range(0, len(tab), 2):
        addr += tab[k]
        delta = tab[k + 1]
        line += delta - 256 if delta >= 128 else delta
        out[addr] = line
    return out


def disassemble(co: Code36, out, indent=""):
    out.write(f"{indent}Disassembly of {co!r}: args={co.varnames[:co.argcount]} locals={co.varnames[co.argcount:]}\n")
    lines = line_starts(co)
    code, ext, targets = co.code, 0, set()
    for i in range(0, len(code), 2):          # jump targets first
        op, arg = code[i], code[i + 1] | ext
        ext = (arg << 8) if op == 144 else 0
        if op in HASJREL:
            targets.add(i + 2 + arg)
        elif op in HASJABS:
            targets.add(arg)
    ext = 0
    for i in range(0, len(code), 2):
        op, arg = code[i], code[i + 1] | ext
        ext = (arg << 8) if op == 144 else 0
        name = OPNAMES.get(op, f"<{op}>")
        note = ""
        if op >= HAVE_ARGUMENT:
            if op in HASCONST:
                note = repr(co.consts[arg])
            elif op in HASNAME:
                note = co.names[arg]
            elif op in HASLOCAL:
                note = co.varnames[arg]
            elif op in HASFREE:
                cells = co.cellvars + co.freevars
                note = cells[arg] if arg < len(cells) else "?"
            elif op in HASJREL:
                note = f"to {i + 2 + arg}"
            elif op in HASJABS:
                note = f"to {arg}"
            elif op == 107:
                note = CMP_OP[arg] if arg < len(CMP_OP) else "?"
        src = f"{lines[i]:>5}" if i in lines else "     "
        mark = ">>" if i in targets else "  "
        argtxt = f"{arg:>5}" if op >= HAVE_ARGUMENT else "     "
        out.write(f"{indent}{src} {mark} {i:>5} {name:<28}{argtxt} {('(' + note + ')') if note else ''}\n")
    out.write("\n")
    for const in co.consts:
        if isinstance(const, Code36):
            disassemble(const, out, indent)


def load_pyc(path: str) -> Code36:
    data = open(path, "rb").read()
    if data[:2] != b"\x33\x0d":
        raise ValueError("not a CPython 3.6 pyc (magic 3379)")
    return Reader(data[12:]).load()     # 3.6 header: magic, mtime, source size


def main(argv):
    co = load_pyc(argv[1])
    sys.stdout.write(f"# CPython 3.6 bytecode of {argv[1].rsplit('/', 1)[-1]} (source file recorded: {co.filename})\n\n")
    disassemble(co, sys.stdout)


if __name__ == "__main__":
    main(sys.argv)
