"""Expression tapes for mcm_chain (include/mcm_b200.h): trace a Python expression
over symbolic particles into the 3-address tape of the device register machine.

In the reference, `sin(a)*a^2 + exp(-a)` on Particles is five broadcasts
(src/register_primitive.jl:44-58,94-97); Julia users fuse them with broadcast
syntax.  Here `fused(f)(a, b, ...)` records f once on symbolic operands, allocates
the 16 value slots by liveness, and issues ONE mcm_chain launch.
"""
from __future__ import annotations

import ctypes as C
import numbers
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L

E = L.ENUMS
NSLOTS = E["MCM_TAPE_SLOTS"]
CONST = E["MCM_TAPE_CONST"]
T_UNARY, T_BINARY, T_TERNARY = E["MCM_TAPE_UNARY"], E["MCM_TAPE_BINARY"], E["MCM_TAPE_TERNARY"]
T_LOAD, T_STORE, T_MOV, T_POWI = E["MCM_TAPE_LOAD"], E["MCM_TAPE_STORE"], E["MCM_TAPE_MOV"], E["MCM_TAPE_POWI"]


class TapeError(L.ArgumentError):
    def __init__(self, msg):
        L.MCMError.__init__(self, L.MCM_ERR_INVALID, msg)


_counter = [0]


class Expr:
    """A symbolic particle value (node of the expression DAG)."""

    __slots__ = ("kind", "op", "args", "imm", "order", "index")
    __array_priority__ = 1000

    def __init__(self, kind, op=None, args=(), imm=0, index=0):
        self.kind, self.op, self.args, self.imm, self.index = kind, op, tuple(args), imm, index
        _counter[0] += 1
        self.order = _counter[0]

    # -- arithmetic (register_primitive_multi list, src/particles.jl:183) --
    def __add__(self, o): return _bin("ADD", self, o)
    def __radd__(self, o): return _bin("ADD", o, self)
    def __sub__(self, o): return _bin("SUB", self, o)
    def __rsub__(self, o): return _bin("SUB", o, self)
    def __mul__(self, o): return _bin("MUL", self, o)
    def __rmul__(self, o): return _bin("MUL", o, self)
    def __truediv__(self, o): return _bin("DIV", self, o)
    def __rtruediv__(self, o): return _bin("DIV", o, self)
    def __mod__(self, o): return _bin("MOD", self, o)
    def __neg__(self): return _un("NEG", self)
    def __pos__(self): return _un("PLUS", self)
    def __abs__(self): return _un("ABS", self)

    def __pow__(self, o):
        # p^i::Integer -> literal_pow / integer power (src/particles.jl:383-391)
        if isinstance(o, numbers.Integral) and not isinstance(o, bool):
            k = int(o)
            if k == 2: return _un("SQR", self)
            if k == 3: return _un("CUBE", self)
            if k == -1: return _un("INV", self)
            if k == 1: return _un("PLUS", self)
            if -128 <= k <= 127: return Expr("op", T_POWI, (self,), imm=k)
        return _bin("POW", self, o)

    def __rpow__(self, o): return _bin("POW", o, self)


def _as_expr(x):
    if isinstance(x, Expr):
        return x
    if isinstance(x, numbers.Real):
        return Expr("const", imm=float(x))
    raise TapeError(f"cannot trace operand of type {type(x).__name__}")


def _un(name, a):
    return Expr("op", T_UNARY + E["MCM_U_" + name], (_as_expr(a),))


def _bin(name, a, b):
    return Expr("op", T_BINARY + E["MCM_B_" + name], (_as_expr(a), _as_expr(b)))


def _ter(name, a, b, c):
    return Expr("op", T_TERNARY + E["MCM_T_" + name], (_as_expr(a), _as_expr(b), _as_expr(c)))


@dataclass
class Tape:
    """A compiled tape: instruction list + constant table + loop section."""
    ins: list = field(default_factory=list)        # (op, dst, a, b, c)
    consts: list = field(default_factory=list)
    nin: int = 0
    nout: int = 0
    loop_begin: int = -1
    loop_end: int = -1
    out_slots: list = field(default_factory=list)  # slot holding each output at the end (for stats)

    def const(self, v: float) -> int:
        v = float(v)
        for i, c in enumerate(self.consts):
            if c == v and np.signbit(c) == np.signbit(v):
                return CONST | i
        if len(self.consts) >= E["MCM_TAPE_MAX_CONST"]:
            raise TapeError("too many distinct constants in one fused chain (max 32)")
        self.consts.append(v)
        return CONST | (len(self.consts) - 1)

    def emit(self, op, dst, a=0, b=0, c=0):
        if len(self.ins) >= E["MCM_TAPE_MAX_INS"]:
            raise TapeError("fused chain too long (max %d instructions)" % E["MCM_TAPE_MAX_INS"])
        self.ins.append((int(op), int(dst), int(a), int(b) & 0xFF, int(c)))

    def as_ctypes(self):
        arr = (L.TapeIns * max(1, len(self.ins)))()
        for i, (op, d, a, b, c) in enumerate(self.ins):
            arr[i].op, arr[i].dst, arr[i].a, arr[i].b, arr[i].c = op, d, a, b, c
        return arr

    @property
    def bounds(self):
        n = len(self.ins)
        lb = n if self.loop_begin < 0 else self.loop_begin
        le = n if self.loop_end < 0 else self.loop_end
        return lb, le


def _toposort(outputs):
    seen, order = set(), []

    def visit(e):
        if id(e) in seen:
            return
        seen.add(id(e))
        for a in e.args:
            visit(a)
        order.append(e)

    for o in outputs:
        visit(o)
    order.sort(key=lambda e: e.order)  # creation order == Python evaluation order (deterministic tapes)
    return order


def compile_exprs(outputs, nin: int) -> Tape:
    """Straight-line tape computing `outputs` (Exprs) from inputs 0..nin-1."""
    outputs = [_as_expr(o) for o in outputs]
    order = _toposort(outputs)
    last_use = {}
    for pos, e in enumerate(order):
        for a in e.args:
            last_use[id(a)] = pos
    for o in outputs:
        last_use[id(o)] = len(order) + 1  # live until stored
    t = Tape(nin=nin, nout=len(outputs))
    free = list(range(NSLOTS))
    loc = {}

    def alloc():
        if not free:
            raise TapeError("expression needs more than %d live values; split it" % NSLOTS)
        return free.pop(0)

    def operand(e):
        return t.const(e.imm) if e.kind == "const" else loc[id(e)]

    for pos, e in enumerate(order):
        if e.kind == "const":
            continue
        if e.kind == "in":
            s = alloc()
            t.emit(T_LOAD, s, e.index)
            loc[id(e)] = s
        else:
            ops = [operand(a) for a in e.args]
            # free operands whose last use is this instruction BEFORE allocating dst (in-place reuse)
            for a in e.args:
                if a.kind != "const" and last_use.get(id(a)) == pos and id(a) in loc and loc[id(a)] not in free:
                    free.append(loc[id(a)])
            free.sort()
            s = alloc()
            if e.op == T_POWI:
                t.emit(T_POWI, s, ops[0], e.imm)
            else:
                t.emit(e.op, s, *ops)
            loc[id(e)] = s
        if last_use.get(id(e)) is None and e.kind != "const":  # dead value
            free.append(loc[id(e)]); free.sort()
    for k, o in enumerate(outputs):
        if o.kind == "const":
            s = alloc()
            t.emit(T_MOV, s, t.const(o.imm))
            loc[id(o)] = s
        t.emit(T_STORE, k, loc[id(o)])
        t.out_slots.append(loc[id(o)])
    return t


def trace(fn, nin: int) -> Tape:
    ins = [Expr("in", index=i) for i in range(nin)]
    out = fn(*ins)
    outs = list(out) if isinstance(out, (tuple, list)) else [out]
    return compile_exprs(outs, nin)


def trace_loop(step, nstate: int, nparam: int, prologue=None) -> Tape:
    """Tape for `state = step(*state, *params)` iterated `reps` times in registers.

    inputs 0..nstate-1 are the initial state, nstate.. the (uncertain) parameters;
    outputs 0..nstate-1 the final state.  `prologue(*params)` may derive loop-invariant
    values once (returned tuple replaces params inside the loop).
    """
    t = Tape(nin=nstate + nparam, nout=nstate)
    # fixed slots: state 0..nstate-1, loop-invariant params after that
    for i in range(nstate):
        t.emit(T_LOAD, i, i)
    inv_exprs = [Expr("in", index=nstate + i) for i in range(nparam)]
    if prologue is not None:
        derived = prologue(*inv_exprs)
        derived = list(derived) if isinstance(derived, (tuple, list)) else [derived]
    else:
        derived = inv_exprs
    ninv = len(derived)
    if nstate + ninv >= NSLOTS:
        raise TapeError("too many state + parameter values for %d slots" % NSLOTS)
    reserved = nstate + ninv

    def lower(outputs, fixed, first_free, pinned):
        """emit code for `outputs`; `fixed` maps id(Expr)->slot for already-resident values."""
        order = _toposort(outputs)
        last_use = {}
        for pos, e in enumerate(order):
            for a in e.args:
                last_use[id(a)] = pos
        for o in outputs:
            last_use[id(o)] = len(order) + 1
        free = list(range(first_free, NSLOTS))
        loc = dict(fixed)
        for pos, e in enumerate(order):
            if e.kind == "const" or id(e) in fixed:
                continue
            if e.kind == "in":
                if not free: raise TapeError("out of slots")
                s = free.pop(0)
                t.emit(T_LOAD, s, e.index)
                loc[id(e)] = s
                continue
            ops = [t.const(a.imm) if a.kind == "const" else loc[id(a)] for a in e.args]
            for a in e.args:
                if a.kind != "const" and last_use.get(id(a)) == pos and loc.get(id(a), -1) >= first_free \
                        and loc[id(a)] not in free and id(a) not in pinned:
                    free.append(loc[id(a)])
            free.sort()
            if not free: raise TapeError("loop body needs more than %d live values" % NSLOTS)
            s = free.pop(0)
            if e.op == T_POWI: t.emit(T_POWI, s, ops[0], e.imm)
            else: t.emit(e.op, s, *ops)
            loc[id(e)] = s
        return loc

    # prologue: derived invariants into slots nstate..nstate+ninv-1
    loc = lower(derived, {}, reserved, set())
    for i, d in enumerate(derived):
        src = t.const(d.imm) if d.kind == "const" else loc[id(d)]
        t.emit(T_MOV, nstate + i, src)
    t.loop_begin = len(t.ins)
    st = [Expr("in", index=i) for i in range(nstate)]
    pv = [Expr("in", index=nstate + i) for i in range(ninv)]
    fixed = {id(e): i for i, e in enumerate(st)}
    fixed.update({id(e): nstate + i for i, e in enumerate(pv)})
    new = step(*st, *pv)
    new = [_as_expr(x) for x in (list(new) if isinstance(new, (tuple, list)) else [new])]
    if len(new) != nstate:
        raise TapeError("step must return %d state values" % nstate)
    loc = lower(new, fixed, reserved, {id(x) for x in new})
    # The write-back is a PARALLEL assignment state <- new: a step may return an old state value in another
    # position ((a, b) -> (a + b, a), (a, b) -> (b, a)), so a move may not overwrite a slot that a later move
    # still reads.  Emit moves whose destination no pending move reads; break cycles through a spare slot.
    pending = {}
    for i, x in enumerate(new):
        src = t.const(x.imm) if x.kind == "const" else loc[id(x)]
        if src != i:
            pending[i] = src
    while pending:
        ready = [d for d in pending if d not in pending.values()]
        if ready:
            for d in ready:
                t.emit(T_MOV, d, pending.pop(d))
            continue
        # only cycles are left: park one source in a slot nobody reads or writes any more
        busy = set(pending) | set(pending.values()) | set(range(reserved))
        spare = [s for s in range(reserved, NSLOTS) if s not in busy]
        if not spare:
            raise TapeError("no spare slot to resolve a cyclic state update")
        d = next(iter(pending))
        t.emit(T_MOV, spare[0], d)
        for k, s in pending.items():
            if s == d:
                pending[k] = spare[0]
    t.loop_end = len(t.ins)
    for i in range(nstate):
        t.emit(T_STORE, i, i)
        t.out_slots.append(i)
    return t


def run_tape(t: Tape, inputs, n: int, dtype: int, outputs=None, stats_output: int | None = None, reps: int = 0,
             sharded: bool = False, wait: bool = True):
    """Launch mcm_chain (or mcm_chain_sharded: statistics of the GLOBAL particle set, merged across ranks inside the
    kernel).  inputs/outputs: device pointers (ints / c_void_p).  Returns the six moments {n, mean, M2, min, max,
    sum} of output `stats_output` if requested, else None.  wait=False returns the launch's TICKET instead and does
    not synchronise: collect with stats_wait(ticket) (any of the last 8 launches)."""
    lib = L.lib()
    if len(inputs) != t.nin:
        raise TapeError(f"tape expects {t.nin} inputs, got {len(inputs)}")
    ins = (C.c_void_p * max(1, t.nin))(*[C.c_void_p(int(p)) for p in inputs])
    outs_list = list(outputs) if outputs is not None else [None] * t.nout
    outs = (C.c_void_p * max(1, t.nout))(*[C.c_void_p(int(p)) if p else C.c_void_p(0) for p in outs_list])
    consts = np.asarray(t.consts if t.consts else [0.0], dtype=np.float64)
    lb, le = t.bounds
    stats = np.zeros(6)
    slot = -1 if stats_output is None else t.out_slots[stats_output]
    fn = lib.mcm_chain_sharded if sharded else lib.mcm_chain
    L.check(fn(dtype, t.as_ctypes(), len(t.ins), lb, le, int(reps), consts.ctypes.data_as(L._dp),
               len(t.consts), ins, t.nin, outs, t.nout, int(n), slot,
               stats.ctypes.data_as(L._dp) if (slot >= 0 and wait) else None))
    if slot < 0:
        return None
    return stats if wait else int(lib.mcm_chain_ticket())


def stats_wait(ticket: int) -> np.ndarray:
    """Statistics of the launch with this ticket (run_tape(..., wait=False)); waits for that launch only."""
    stats = np.zeros(6)
    L.check(L.lib().mcm_chain_stats_wait(int(ticket), stats.ctypes.data_as(L._dp)))
    return stats
