"""Run-time compilation of user vector fields.

The reference hands every template to PyRates, which generates the right-hand side `func(t, y, *args)` as torch code at run time
(rectipy/nodes.py:232-262: `_circuit_from_yaml` -> `CircuitTemplate.get_run_func`).  The engine ships compiled vector fields for
the templates the reference ships; an operator whose equations match none of them takes this route instead:

    YAML equations --sympy--> f_k(y, u, I, spike, params), the Jacobians df/dy, df/du, df/dI, df/dspike, df/dparams and the source
    expression + its derivative --C code printer--> one CUDA translation unit (forward Euler step, reverse-time adjoint step,
    source initialisation) against the argument records of csrc/rp_jit_abi.cuh --NVRTC (sm_100a)--> cubin -->
    rp_plan_set_jit_module (C ABI), after which rp_forward / rp_backward launch these kernels inside their per-step loops in place
    of k_fwd_step<MODEL> / k_adj_step<MODEL>.  The contractions (W.src, W^T.g, g (x) src), the Observer, the checkpoints and the
    autograd plumbing are the engine's own.

Semantics are those of the reference's node classes (rectipy/nodes.py:166-170,382-392,451-465,468-481): explicit Euler; for spiking
nodes spike = heaviside(v - theta) with the surrogate gradient, the spike enters the field as spike/dt, the reset variable is blended
with the detached spike; outputs are pre-update slices -- except MultiSpikeResetNet (several spike / reset variable pairs), whose
outputs and recorded variables are post-update.  Restrictions (checked, NotImplementedError otherwise): first-order explicit
equations, at most RP_MAX_SV state variables, df/du must not depend on u, I or the spike; no mean-field `mean(.)` terms.
"""
from __future__ import annotations

import hashlib
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from . import _cabi as abi

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
_INCLUDE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")


@dataclass
class JitField:
    """What `templates._spec_from_ops` keeps of an operator set that matches no compiled field."""
    states: List[str]                       # state variables in equation order (short names)
    rhs: Dict[str, object]                  # state -> sympy expression (algebraic variables substituted)
    algebraic: Dict[str, object]            # algebraic variable -> sympy expression of states / params / inputs
    params: List[str]                       # parameter names (short), in slot order
    inputs: List[str]                       # `input`-typed variable names
    defaults: Dict[str, float] = field(default_factory=dict)
    owner: Dict[str, str] = field(default_factory=dict)     # short name -> operator that defines it ("op/name" keys of the spec)
    ops: Tuple[str, ...] = ()


@dataclass
class JitProgram:
    source: str
    image: bytes
    key: str
    nsv: int
    spiking: int                            # number of spike variables (planes 0..spiking-1 are thresholded and reset)
    src_plane: int
    planes: Dict[str, int]                  # state variable -> engine plane
    param_slots: Dict[str, int]             # parameter -> ABI slot
    post_out: bool = False                  # MultiSpikeResetNet: outputs / recorded variables are post-update


_PROGRAMS: Dict[str, JitProgram] = {}


def program(key: str) -> JitProgram:
    return _PROGRAMS[key]


def _sym(name):
    import sympy
    return sympy.Symbol(name)


def parse_field(ops) -> JitField:
    """Operator definitions (templates.OperatorDef) -> symbolic field.  Operators of one node share a namespace of short variable
    names, as PyRates connects them (li_op's `v` is tanh_op's input `v`)."""
    import sympy
    from .templates import _rhs_expr
    diff_eqs, alg_eqs = {}, {}
    variables: Dict[str, object] = {}
    owner: Dict[str, str] = {}
    for op in ops:
        for name, val in op.variables.items():
            if name in variables and isinstance(variables[name], tuple) and variables[name][0] in ("output", "variable"):
                continue                     # an earlier operator defines it (output there, input here)
            variables[name] = val
            owner[name] = op.name
        for eq in op.equations:
            lhs, _, rhs = eq.partition("=")
            lhs, rhs = lhs.strip(), rhs.strip()
            if not rhs:
                raise NotImplementedError(f"rectipy_b200.jit: cannot parse equation {eq!r}")
            if "mean(" in rhs.replace(" ", ""):
                raise NotImplementedError("rectipy_b200.jit: population-mean terms (mean(.)) are only available in the compiled iku / ik_biexp fields")
            m = re.match(r"^([A-Za-z_][A-Za-z_0-9]*)'$", lhs) or re.match(r"^d/dt\s*\*?\s*([A-Za-z_][A-Za-z_0-9]*)$", lhs)
            if m:
                diff_eqs[m.group(1)] = _rhs_expr(rhs)
            elif re.match(r"^[A-Za-z_][A-Za-z_0-9]*$", lhs):
                alg_eqs[lhs] = _rhs_expr(rhs)
            else:
                raise NotImplementedError(f"rectipy_b200.jit: unsupported left-hand side {lhs!r} (first-order explicit equations only)")
    if not diff_eqs:
        raise NotImplementedError("rectipy_b200.jit: the operator set has no differential equation")
    if len(diff_eqs) > abi.RP_MAX_SV:
        raise NotImplementedError(f"rectipy_b200.jit: {len(diff_eqs)} state variables (max {abi.RP_MAX_SV})")
    # substitute algebraic variables (possibly chained) into everything
    for _ in range(len(alg_eqs) + 1):
        alg_eqs = {k: e.subs({_sym(a): alg_eqs[a] for a in alg_eqs if a != k}) for k, e in alg_eqs.items()}
    rhs = {k: e.subs({_sym(a): alg_eqs[a] for a in alg_eqs}) for k, e in diff_eqs.items()}
    states = list(diff_eqs.keys())
    used = set().union(*[{str(s) for s in e.free_symbols} for e in list(rhs.values()) + list(alg_eqs.values())])
    inputs = [n for n, v in variables.items() if isinstance(v, tuple) and v[0] == "input" and n not in states and n not in alg_eqs]
    params = [n for n in sorted(used) if n not in states and n not in inputs and n not in alg_eqs]
    for n in params:
        if n not in variables or isinstance(variables[n], tuple):
            raise NotImplementedError(f"rectipy_b200.jit: symbol {n!r} is neither a state variable, an input nor a parameter with a default value")
    if len(params) > abi.RP_NUM_PARAMS - 1:
        raise NotImplementedError(f"rectipy_b200.jit: {len(params)} parameters (max {abi.RP_NUM_PARAMS - 1})")
    defaults = {n: (float(v[1]) if isinstance(v, tuple) else float(v)) for n, v in variables.items()}
    return JitField(states=states, rhs=rhs, algebraic=alg_eqs, params=params, inputs=inputs, defaults=defaults, owner=owner,
                    ops=tuple(op.name for op in ops))


def unbound_spec(ops):
    """TemplateSpec of an operator set that matches no compiled field, before the node's variable roles are known."""
    from .templates import TemplateSpec
    fld = parse_field(ops)
    key = lambda n: f"{fld.owner.get(n, fld.ops[0])}/{n}"
    return TemplateSpec(name="jit", model=abi.RP_JIT, ops=fld.ops, state_vars=[(key(s), fld.defaults.get(s, 0.0)) for s in fld.states],
                        params={key(p): (slot, fld.defaults[p]) for p, slot in param_slot_table(fld.params).items()},
                        source_var="", target_var="", input_vars={key(n): 0 for n in fld.inputs}, spike_var=None,
                        out_vars={}, planes={}, jit_field=fld)


def bind_spec(spec, source_var, target_var, input_var, spike_var, reset_var):
    """Fix the roles of the template's variables (what `from_pyrates` receives, rectipy/nodes.py:112-164,363-380,438-449), generate
    and compile the kernels, and return the complete TemplateSpec (with `jit_program`).  `spike_var` / `reset_var` given as lists
    select the MultiSpikeResetNet semantics (several spike variables, post-update outputs)."""
    from .templates import TemplateSpec
    fld: JitField = spec.jit_field
    short = lambda n: None if n is None else [q for q in str(n).split("/") if q][-1]
    key = lambda n: f"{fld.owner.get(n, fld.ops[0])}/{n}"
    multi = isinstance(spike_var, (list, tuple))
    spks = [short(v) for v in spike_var] if multi else ([short(spike_var)] if spike_var is not None else [])
    rsts = [short(v) for v in reset_var] if isinstance(reset_var, (list, tuple)) else ([short(reset_var)] if spks else [])
    if len(spks) != len(rsts):
        raise ValueError("`spike_var` and `reset_var` must name the same number of variables")
    src, tgt, inp = short(source_var), short(target_var), short(input_var)
    if src is None:
        src = fld.states[0]
    if src not in fld.states and src not in fld.algebraic:
        raise KeyError(f"Variable {source_var} was not found on the node template.")
    for n in [tgt, inp] + spks:
        if n is not None and n not in fld.inputs:
            raise KeyError(f"Variable {n} was not found among the template's input variables.")
    for r in rsts:
        if r not in fld.states:
            raise KeyError(f"Variable {r} was not found on the node template.")
    prog = build_program(fld, src, tgt, inp, spks, rsts, post_out=multi)
    params = {key(p): (slot, fld.defaults[p]) for p, slot in prog.param_slots.items()}
    params[ONE_KEY] = (abi.RP_P_K, 1.0)
    planes = {key(s): pl for s, pl in prog.planes.items()}
    return TemplateSpec(name="jit:" + prog.key[:10], model=abi.RP_JIT, ops=fld.ops, state_vars=list(spec.state_vars), params=params,
                        source_var=key(src), target_var=key(tgt) if tgt else "", input_vars={key(inp): 0} if inp else {},
                        spike_var=key(spks[0]) if spks else None, out_vars={k: pl for k, pl in planes.items() if pl < 3}, planes=planes,
                        jit_field=fld, jit_program=prog)


#: the engine multiplies the recurrent weights by parameter slot RP_P_K; a run-time compiled field keeps its coupling constants in
#: the equations, so the node carries a constant 1 there
ONE_KEY = "jit/one"


def param_slot_table(params: List[str]) -> Dict[str, int]:
    """Parameter -> ABI slot.  Slot RP_P_K stays free: the engine folds `params[RP_P_K]` into the weights, and a run-time compiled
    field keeps its coupling constants inside the equations -- the node passes a constant 1 there."""
    slots = [q for q in range(abi.RP_NUM_PARAMS) if q != abi.RP_P_K]
    return {n: slots[i] for i, n in enumerate(params)}


_PRINTER = None


def _ccode(expr) -> str:
    """float32 C code of a sympy expression; small integer powers become products (`v^2` of the templates is `v*v`, as PyRates'
    torch code computes it -- not powf)."""
    global _PRINTER
    if _PRINTER is None:
        from sympy.codegen.ast import real, float32
        from sympy.printing.c import C99CodePrinter

        class Printer(C99CodePrinter):
            def _print_Pow(self, e):
                if e.exp.is_Integer and 2 <= abs(int(e.exp)) <= 4:
                    base = self.parenthesize(e.base, 1000)        # atoms stay bare, everything else is bracketed
                    prod = "*".join([base] * abs(int(e.exp)))
                    return f"({prod})" if int(e.exp) > 0 else f"(1.0F/({prod}))"
                return super()._print_Pow(e)
        _PRINTER = Printer({"type_aliases": {real: float32}})
    return _PRINTER.doprint(expr)


def build_program(fld: JitField, source_var: str, target_var: Optional[str], input_var: Optional[str],
                  spike_vars: Sequence[str] = (), reset_vars: Sequence[str] = (), post_out: bool = False) -> JitProgram:
    """Generate, compile (NVRTC, sm_100a) and cache the kernels of one node configuration.  `spike_vars[j]` is the input variable that
    receives spike_j / dt, `reset_vars[j]` the state variable it thresholds and resets (plane j); `post_out`: MultiSpikeResetNet."""
    import sympy
    spike_vars, reset_vars = list(spike_vars), list(reset_vars)
    nspk = len(spike_vars)
    if len(reset_vars) != nspk or len(set(reset_vars)) != nspk:
        raise ValueError("every spike variable needs its own reset variable")
    states = [s for s in reset_vars] + [s for s in fld.states if s not in reset_vars]       # the engine thresholds / resets planes 0..nspk-1
    for r in reset_vars:
        if r not in fld.states:
            raise KeyError(r)
    planes = {s: i for i, s in enumerate(states)}
    nsv = len(states)
    slots = param_slot_table(fld.params)
    # symbols -> C identifiers
    U, I = _sym("rp_u"), _sym("rp_I")
    SPK = [_sym(f"spk{j}") for j in range(nspk)]
    sub = {_sym(s): _sym(f"y{planes[s]}") for s in states}
    sub.update({_sym(p): _sym(f"p{slots[p]}") for p in fld.params})
    if target_var is not None:
        if target_var not in fld.inputs:
            raise KeyError(target_var)
        sub[_sym(target_var)] = U
    for n in fld.inputs:
        if n == target_var:
            continue
        if input_var is not None and n == input_var:
            sub[_sym(n)] = I
        elif n in spike_vars:
            sub[_sym(n)] = SPK[spike_vars.index(n)]
        else:
            sub[_sym(n)] = sympy.Float(fld.defaults.get(n, 0.0))     # an input nobody drives keeps its default
    f = [fld.rhs[s].subs(sub) for s in states]              # not simplified: keep the user's evaluation order
    if source_var in planes:
        src_plane, src = planes[source_var], _sym(f"y{planes[source_var]}")
    elif source_var in fld.algebraic:
        src_plane, src = -1, fld.algebraic[source_var].subs(sub)
    else:
        raise KeyError(source_var)
    ys = [_sym(f"y{k}") for k in range(nsv)]
    J = [[sympy.diff(f[k], ys[l]) for l in range(nsv)] for k in range(nsv)]
    Ju = [sympy.diff(f[k], U) for k in range(nsv)]
    JI = [sympy.diff(f[k], I) for k in range(nsv)]
    Js = [[sympy.diff(f[k], SPK[j]) for j in range(nspk)] for k in range(nsv)]
    for e in Ju:
        if e.free_symbols & ({U, I} | set(SPK)):
            raise NotImplementedError("rectipy_b200.jit: d f / d (coupling input) must not depend on the coupling input, the external input or the spike")
    plist = [(n, slots[n]) for n in fld.params]
    Jp = [[sympy.diff(f[k], _sym(f"p{q}")) for (_, q) in plist] for k in range(nsv)]
    dsrc = [sympy.diff(src, ys[l]) for l in range(nsv)]

    def lines(prefix, exprs):
        return "\n".join(f"    {prefix}[{i}] = {_ccode(e)};" for i, e in enumerate(exprs))

    def lines2(prefix, mat):
        return "\n".join(f"    {prefix}[{i}][{j}] = {_ccode(e)};" for i, row in enumerate(mat) for j, e in enumerate(row))

    load_y = "\n".join(f"    const float y{k} = y[{k}];" for k in range(nsv))
    load_p = "\n".join(f"    const float p{q} = rp::ldp(mp, {q}, i, b);" for (_, q) in plist)
    load_s = "\n".join(f"    const float spk{j} = spk[{j}];" for j in range(nspk))
    source = _TEMPLATE.format(NSV=nsv, NPAR=max(1, len(plist)), NPAR_REAL=len(plist), NSPK=nspk, NSPK1=max(1, nspk), POST_OUT=int(post_out),
                              LOAD_Y=load_y, LOAD_P=load_p, LOAD_S=load_s,
                              F=lines("f", f), J=lines2("J", J), JU=lines("Ju", Ju), JI=lines("JI", JI), JS=lines2("Js", Js) if nspk else "",
                              JP=lines2("Jp", Jp) if plist else "", DSRC=lines("dsrc", dsrc), SRC=_ccode(src),
                              PSLOTS=", ".join(str(q) for (_, q) in plist) if plist else "0")
    key = hashlib.sha1(source.encode()).hexdigest()
    if key not in _PROGRAMS:
        _PROGRAMS[key] = JitProgram(source=source, image=compile_cuda(source), key=key, nsv=nsv, spiking=nspk, src_plane=src_plane,
                                    planes=planes, param_slots=slots, post_out=bool(post_out))
    return _PROGRAMS[key]


def compile_cuda(source: str) -> bytes:
    """NVRTC: CUDA C++ -> sm_100a cubin.  Works without a GPU (the CPU tests compile, only loading needs a device)."""
    from cuda.bindings import nvrtc
    with open(os.path.join(_CSRC, "rp_jit_abi.cuh")) as fh:
        abi_h = fh.read()
    with open(os.path.join(_INCLUDE, "rectipy_b200.h")) as fh:
        api_h = fh.read()
    err, prog = nvrtc.nvrtcCreateProgram(source.encode(), b"rp_jit_field.cu", 2, [abi_h.encode(), api_h.encode()],
                                         [b"rp_jit_abi.cuh", b"rectipy_b200.h"])
    if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        raise RuntimeError(f"nvrtcCreateProgram failed: {err}")
    opts = [b"--gpu-architecture=sm_100a", b"-lineinfo", b"--std=c++17", b"--fmad=true"]
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, logn = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * logn
    nvrtc.nvrtcGetProgramLog(prog, log)
    if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        raise RuntimeError("rectipy_b200.jit: NVRTC could not compile the generated vector field:\n" + log.decode(errors="replace")
                           + "\n--- generated source ---\n" + source)
    _, n = nvrtc.nvrtcGetCUBINSize(prog)
    image = b" " * n
    nvrtc.nvrtcGetCUBIN(prog, image)
    nvrtc.nvrtcDestroyProgram(prog)
    return bytes(image)


_TEMPLATE = r'''// generated by rectipy_b200/jit.py from a user template -- forward Euler step, reverse-time adjoint step, source initialisation
#include "rectipy_b200.h"
#include "rp_jit_abi.cuh"

#define NSV {NSV}
#define NPAR {NPAR}
#define NPAR_REAL {NPAR_REAL}
#define NSPK {NSPK}          // spike variables: planes 0..NSPK-1 are thresholded and reset
#define NSPK1 {NSPK1}
#define POST_OUT {POST_OUT}  // 1: MultiSpikeResetNet (outputs are post-update slices)
__device__ const int PSLOT[NPAR] = {{ {PSLOTS} }};

// right-hand side f_k(y, u, I, spike_j / dt)
__device__ __forceinline__ void jit_field(const rp::ModelParams& mp, int i, int b, const float* y, float rp_u, float rp_I, const float* spk, float* f) {{
{LOAD_Y}
{LOAD_P}
{LOAD_S}
{F}
}}
__device__ __forceinline__ float jit_src(const rp::ModelParams& mp, int i, int b, const float* y) {{
{LOAD_Y}
{LOAD_P}
    return {SRC};
}}
// Jacobians at (y, u, I, spike arguments)
__device__ __forceinline__ void jit_jac(const rp::ModelParams& mp, int i, int b, const float* y, float rp_u, float rp_I, const float* spk,
                                        float (*J)[NSV], float* Ju, float* JI, float (*Js)[NSPK1], float (*Jp)[NPAR], float* dsrc) {{
{LOAD_Y}
{LOAD_P}
{LOAD_S}
{J}
{JU}
{JI}
{JS}
{JP}
{DSRC}
}}
// spike_j = heaviside(y_j - theta, 1.0); the field receives spike_j / dt (rectipy/nodes.py:383-385,453-456)
__device__ __forceinline__ void jit_spikes(const float* y, float theta, float dt, bool* sp, float* spk) {{
    spk[0] = 0.f;
#pragma unroll
    for (int j = 0; j < NSPK; ++j) {{ sp[j] = y[j] >= theta; spk[j] = sp[j] ? 1.0f / dt : 0.0f; }}
}}
// one Euler step with the reset (the adjoint of a post-update output needs y_(t+1) again)
__device__ __forceinline__ void jit_step(const rp::ModelParams& mp, int i, int b, const float* y, float u, float Iin, float theta, float v_reset,
                                         float dt, float* y1) {{
    bool sp[NSPK1]; float spk[NSPK1], f[NSV];
    jit_spikes(y, theta, dt, sp, spk);
    jit_field(mp, i, b, y, u, Iin, spk, f);
#pragma unroll
    for (int k = 0; k < NSV; ++k) y1[k] = y[k] + dt * f[k];
#pragma unroll
    for (int j = 0; j < NSPK; ++j) if (sp[j]) y1[j] = v_reset;          // reset (blend with the detached spike / masked assignment)
}}

struct JitInitSrcArgs {{ int N, B; const float* y; rp::ModelParams mp; float* src; int ld; }};
extern "C" __global__ void __launch_bounds__(256) rp_jit_init_src(JitInitSrcArgs a) {{
    const size_t plane = (size_t)a.B * a.N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < plane; idx += (size_t)gridDim.x * blockDim.x) {{
        const int b = (int)(idx / a.N), i = (int)(idx - (size_t)b * a.N);
        float y[NSV];
#pragma unroll
        for (int k = 0; k < NSV; ++k) y[k] = a.y[(size_t)k * plane + idx];
        a.src[(size_t)b * a.ld + i] = jit_src(a.mp, i, b, y);
    }}
}}

// y_{{t+1}} = y_t + dt f(y_t, u_t, I_t, spike_t / dt), then the reset (rectipy/nodes.py:166-170,382-392,451-465)
extern "C" __global__ void __launch_bounds__(256) rp_jit_fwd_step(rp::FwdStepArgs a) {{
    const size_t plane = (size_t)a.B * a.N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < plane; idx += (size_t)gridDim.x * blockDim.x) {{
        const int b = (int)(idx / a.N), i = (int)(idx - (size_t)b * a.N);
        float y[NSV], y1[NSV];
#pragma unroll
        for (int k = 0; k < NSV; ++k) y[k] = a.y_cur[(size_t)k * plane + idx];
        const float u = a.u[(size_t)b * a.ldu + i];
        const float Iin = rp::input_current(a.in_mode, a.m, a.x_t, a.W_in, a.N, b, i);
        jit_step(a.mp, i, b, y, u, Iin, a.theta, a.v_reset, a.dt, y1);
        if (a.urec_out) a.urec_out[idx] = u;
#pragma unroll
        for (int k = 0; k < NSV; ++k) a.y_next[(size_t)k * plane + idx] = y1[k];
        if (a.src_next) a.src_next[idx] = jit_src(a.mp, i, b, y1);
    }}
}}

// Few trials (B <= 8): the same step with the recurrent contraction inside -- one warp per neuron row forms u_b = W[i][:] . src_b for
// every trial (coalesced row read, trial vectors from L1/L2), a shuffle tree reduces, lane b steps (trial b, neuron i).
extern "C" __global__ void __launch_bounds__(256) rp_jit_fwd_step_rows(rp::JitRowsArgs r) {{
    const rp::FwdStepArgs& a = r.a;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= a.N) return;
    const size_t plane = (size_t)a.B * a.N;
    float acc[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b] = 0.f;
    const float* wrow = r.W + (size_t)i * r.ldw;
    if ((a.N & 3) == 0) {{                                   // 16-byte loads, several chunks in flight per lane (the loop is latency-bound otherwise)
        const float4* w4 = reinterpret_cast<const float4*>(wrow);
        const int n4 = a.N >> 2;
#pragma unroll 4
        for (int j = lane; j < n4; j += 32) {{
            const float4 w = __ldg(w4 + j);
#pragma unroll
            for (int b = 0; b < 8; ++b) if (b < a.B) {{
                const float4 s = *reinterpret_cast<const float4*>(r.src + (size_t)b * a.N + 4 * j);
                acc[b] = fmaf(w.x, s.x, fmaf(w.y, s.y, fmaf(w.z, s.z, fmaf(w.w, s.w, acc[b]))));
            }}
        }}
    }} else {{
#pragma unroll 8
        for (int j = lane; j < a.N; j += 32) {{
            const float w = __ldg(wrow + j);
#pragma unroll
            for (int b = 0; b < 8; ++b) if (b < a.B) acc[b] = fmaf(w, r.src[(size_t)b * a.N + j], acc[b]);
        }}
    }}
#pragma unroll
    for (int b = 0; b < 8; ++b)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
    float u = 0.f;
#pragma unroll
    for (int b = 0; b < 8; ++b) if (lane == b) u = acc[b];
    if (lane < a.B) {{
        const int b = lane;
        const size_t idx = (size_t)b * a.N + i;
        float y[NSV], y1[NSV];
#pragma unroll
        for (int k = 0; k < NSV; ++k) y[k] = a.y_cur[(size_t)k * plane + idx];
        const float Iin = rp::input_current(a.in_mode, a.m, a.x_t, a.W_in, a.N, b, i);
        jit_step(a.mp, i, b, y, u, Iin, a.theta, a.v_reset, a.dt, y1);
        if (a.urec_out) a.urec_out[idx] = u;
#pragma unroll
        for (int k = 0; k < NSV; ++k) a.y_next[(size_t)k * plane + idx] = y1[k];
        if (a.src_next) a.src_next[idx] = jit_src(a.mp, i, b, y1);
    }}
}}

// gradient of the record window into the output variable (+ dW_out): adj[out] += W_out^T e (readout) | e (dense)
__device__ __forceinline__ void jit_out_grad(const rp::AdjArgs& a, const float* e, float e_scale, int b, int i, size_t idx, float yout, float* adj) {{
    if (a.out_mode == RP_OUT_DENSE) {{ adj[a.out_var] += __ldg(e + idx) * e_scale; return; }}
    float ro = 0.f;
    for (int q = 0; q < a.k; ++q) {{
        const float ev = __ldg(e + (size_t)b * a.k + q) * e_scale;
        ro = fmaf(__ldg(a.W_out + (size_t)q * a.N + i), ev, ro);
        if (a.dW_out) atomicAdd(a.dW_out + (size_t)q * a.N + i, ev * yout);
    }}
    adj[a.out_var] += ro;
}}

// reverse-time adjoint of one step (SURVEY Appendix A, generalised to an arbitrary field): "post" finishes step t (needs
// Z_t = W^T g_t), "pre" prepares g_{{t-1}} and the source value of step t-1
extern "C" __global__ void __launch_bounds__(256) rp_jit_adj_step(rp::AdjArgs a) {{
    const size_t plane = (size_t)a.B * a.N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < plane; idx += (size_t)gridDim.x * blockDim.x) {{
        const int b = (int)(idx / a.N), i = (int)(idx - (size_t)b * a.N);
        float adj[NSV];
#pragma unroll
        for (int k = 0; k < NSV; ++k) adj[k] = a.adj[(size_t)k * plane + idx];
        float J[NSV][NSV], Ju[NSV], JI[NSV], Js[NSV][NSPK1], Jp[NSV][NPAR], dsrc[NSV], spk[NSPK1];
        bool sp[NSPK1];
        float y[NSV];
        if (a.do_post || (POST_OUT && a.e_tm1)) {{
#pragma unroll
            for (int k = 0; k < NSV; ++k) y[k] = __ldg(a.y_t + (size_t)k * plane + idx);
        }}
        if (a.do_post) {{
            const float u = a.urec_t ? __ldg(a.urec_t + idx) : 0.f;
            const float Iin = rp::input_current(a.in_mode, a.m, a.x_t, a.W_in, a.N, b, i);
            jit_spikes(y, a.theta, a.dt, sp, spk);
            jit_jac(a.mp, i, b, y, u, Iin, spk, J, Ju, JI, Js, Jp, dsrc);
            float at[NSV];
#pragma unroll
            for (int k = 0; k < NSV; ++k) at[k] = adj[k];
#pragma unroll
            for (int j = 0; j < NSPK; ++j) if (sp[j]) at[j] = 0.f;      // no gradient through the Euler update of a variable that is reset
            const float Z = a.Z[(size_t)b * a.ldz + i];
            float nw[NSV];
#pragma unroll
            for (int l = 0; l < NSV; ++l) {{
                float acc = at[l];
#pragma unroll
                for (int k = 0; k < NSV; ++k) acc = fmaf(a.dt * at[k], J[k][l], acc);
                nw[l] = fmaf(Z, dsrc[l], acc);
            }}
#pragma unroll
            for (int j = 0; j < NSPK; ++j) {{                            // surrogate: d spike_j / d y_j = 1 / (1 + slope |y_j - theta|)^2; f sees spike_j / dt
                const float dd = 1.0f + a.slope * fabsf(y[j] - a.theta);
                float c = 0.f;
#pragma unroll
                for (int k = 0; k < NSV; ++k) c = fmaf(at[k], Js[k][j], c);
                nw[j] += c / (dd * dd);
            }}
            if (!POST_OUT && a.e_t) jit_out_grad(a, a.e_t, a.e_scale, b, i, idx, y[a.out_var], nw);      // pre-update output: y_t[out]
            float dI = 0.f;
#pragma unroll
            for (int k = 0; k < NSV; ++k) dI = fmaf(a.dt * at[k], JI[k], dI);
            if (a.g_x_t) a.g_x_t[idx] = dI;
            if (a.dW_in && a.in_mode == RP_IN_PROJ)
                for (int j = 0; j < a.m; ++j) atomicAdd(a.dW_in + (size_t)i * a.m + j, dI * __ldg(a.x_t + (size_t)b * a.m + j));
#pragma unroll
            for (int q = 0; q < NPAR_REAL; ++q) {{
                float* dst = a.dparams[PSLOT[q]];
                if (dst) {{
                    float c = 0.f;
#pragma unroll
                    for (int k = 0; k < NSV; ++k) c = fmaf(a.dt * at[k], Jp[k][q], c);
                    atomicAdd(dst + i, c);
                }}
            }}
#pragma unroll
            for (int k = 0; k < NSV; ++k) adj[k] = a.zero_after_post ? 0.f : nw[k];
        }}
        if (POST_OUT && a.e_tm1) jit_out_grad(a, a.e_tm1, a.e_scale_tm1, b, i, idx, y[a.out_var], adj);   // post-update output of step t-1: y_t[out]
        if (a.do_post || (POST_OUT && a.e_tm1)) {{
#pragma unroll
            for (int k = 0; k < NSV; ++k) a.adj[(size_t)k * plane + idx] = adj[k];
        }}
        if (a.do_pre) {{
            float ym[NSV];
#pragma unroll
            for (int k = 0; k < NSV; ++k) ym[k] = __ldg(a.y_tm1 + (size_t)k * plane + idx);
            jit_spikes(ym, a.theta, a.dt, sp, spk);
            jit_jac(a.mp, i, b, ym, 0.f, 0.f, spk, J, Ju, JI, Js, Jp, dsrc);       // d f / d u depends on the state only (checked at code generation)
            float g = 0.f;
#pragma unroll
            for (int k = 0; k < NSV; ++k) g = fmaf(a.dt * ((k < NSPK && sp[k < NSPK ? k : 0]) ? 0.f : adj[k]), Ju[k], g);
            if (a.g) a.g[idx] = g;
            if (a.src) a.src[idx] = jit_src(a.mp, i, b, ym);
        }}
    }}
}}
'''
