"""Model front-end: neuron templates -> engine vector-field id + parameter/variable tables.

The reference hands the YAML template to PyRates, which vectorises N identical nodes and emits a torch function
(rectipy/nodes.py:112-164,232-262).  PyRates is third-party and not part of the hot path; here the operator
templates shipped with the reference (neuron_model_templates/rate_neurons/leaky_integrator.yaml,
spiking_neurons/qif.yaml, spiking_neurons/lif.yaml) are recognised by their equations and mapped onto the CUDA
vector fields of rectipy_b200/csrc/rp_kernels.cuh.  User YAML files in the same PyRates template syntax
(`base:`, `equations: replace/add`, `variables:`) are parsed and matched against the compiled fields -- by equation text
first, then symbolically (sympy: `v*v` for `v^2`, reordered terms, `-(1/tau)*v` for `-v/tau` are the same field) -- so changed
default values and rewritten but identical equations are honoured; an operator set whose equations match none of the compiled
fields is handed to rectipy_b200/jit.py, which generates its CUDA kernels (step, adjoint) with sympy and compiles them with NVRTC
(there is no CPU fallback; what the generator cannot express raises NotImplementedError).
"""
from __future__ import annotations

import importlib.util
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from . import _cabi as abi


@dataclass
class OperatorDef:
    name: str
    equations: List[str]
    variables: Dict[str, object]      # name -> default (float) | ("input", float) | ("output"|"variable", float)


@dataclass
class TemplateSpec:
    name: str                               # engine-side template name: li_tanh, li_sigmoid, qif, qif_sfa, lif
    model: int                              # abi.RP_*
    ops: Tuple[str, ...]                    # operator names as the user sees them (may be renamed in user YAML)
    state_vars: List[Tuple[str, float]]     # ordered ("op/var", initial value); order = equation order
    params: Dict[str, Tuple[int, float]]    # "op/name" -> (abi slot, default)
    source_var: str                         # "op/var" projected by the recurrent weights
    target_var: str                         # "op/var" receiving  W @ source
    input_vars: Dict[str, int]              # "op/var" -> in_target code
    spike_var: Optional[str]                # "op/spike" for spiking templates
    out_vars: Dict[str, int] = field(default_factory=dict)   # "op/var" -> abi.RP_VAR_*
    planes: Dict[str, int] = field(default_factory=dict)     # "op/var" -> engine state plane (0 v, 1 s, 2 x/u); the order of
                                                             # `state_vars` is the reference's y order (equation order)
    fold_param: str = ""                                     # "op/name" of the coupling constant folded into the weights
    jit_field: object = None                                 # rectipy_b200.jit.JitField: equations that match no compiled field
    jit_program: object = None                               # rectipy_b200.jit.JitProgram once the variable roles are bound

    def plane_of(self, key: str) -> int:
        return self.planes[key] if self.planes else [k for k, _ in self.state_vars].index(key)

    @property
    def plane_order(self) -> List[int]:
        """engine plane of each state variable in reference order"""
        return [self.plane_of(k) for k, _ in self.state_vars]

    @property
    def spiking(self) -> bool:
        return self.spike_var is not None

    @property
    def n_sv(self) -> int:
        return len(self.state_vars)

    def resolve(self, name: str, table) -> str:
        """Map a user variable name ("v", "li_op/v", "all/li_op/v", "n0/li_op/v") onto a key of `table`."""
        parts = [p for p in name.split("/") if p]
        if len(parts) >= 2:
            key = "/".join(parts[-2:])
            if key in table:
                return key
        short = parts[-1]
        hits = [k for k in table if k.split("/")[-1] == short]
        if len(hits) == 1:
            return hits[0]
        if len(hits) > 1 and len(parts) >= 2:
            for k in hits:
                if k.split("/")[0] == parts[-2]:
                    return k
        if len(hits) > 1:
            return hits[0]
        raise KeyError(name)


def _canon(eq: str) -> str:
    return re.sub(r"\s+", "", eq)


# canonical equation sets of the compiled vector fields (text of the reference YAML files)
_LI = [_canon("v' = -v/tau + k*r_in + I_ext + eta")]
_TANH = [_canon("r = tanh(v)")]
_SIGMOID = [_canon("r = r_max / (1 + exp(s*(v0-v)))")]
_QIF = [_canon("v' = (v^2 + eta + I_ext)/tau + k*s_in"), _canon("s' = -s/tau_s + spike")]
_QIF_SFA = [_canon("v' = (v^2 + eta - x + I_ext)/tau + k*s_in"), _canon("s' = -s/tau_s + spike"),
            _canon("x' = -x/tau_x + alpha*spike")]
_LIF = [_canon("v' = -v/tau + k*s_in + I_ext + eta"), _canon("s' = -s/tau_s + spike + s_ext")]
_IK = [_canon("v' = (k*(v-v_r)*(v-v_theta) - u + I_ext + eta + g*s_in*(E_r - v)) / C"),
       _canon("u' = (b*(v-v_r) - u) / tau_u + kappa*spike"), _canon("s' = -s/tau_s + spike")]

_IKU = [_IK[0], _canon("u' = (b*(mean(v)-v_r) - u) / tau_u + kappa*mean(spike)"), _IK[2]]
_IK_BIEXP = [_IK[0], _IKU[1], _canon("s' = -s/tau_d + x"), _canon("x' = -x/tau_r + spike")]

_KNOWN_FIELDS = [_LI, _TANH, _SIGMOID, _QIF, _QIF_SFA, _LIF, _IK, _IKU, _IK_BIEXP]
_FUNCS = ("tanh", "exp", "mean", "sigmoid", "sin", "cos", "sqrt", "log", "abs")


def _rhs_expr(rhs: str):
    """sympy expression of an equation's right-hand side; every identifier that is not a function name becomes a plain Symbol
    (so that `E_r`, `I_ext`, `S`, `N`, `beta` ... are never taken for sympy constants / functions)."""
    import sympy
    names = set(re.findall(r"[A-Za-z_][A-Za-z_0-9]*", rhs))
    local = {n: sympy.Symbol(n) for n in names if n not in _FUNCS}
    local["mean"] = sympy.Function("mean")
    return sympy.sympify(rhs.replace("^", "**"), locals=local)


def _same_equation(a: str, b: str) -> bool:
    """Same left-hand side and algebraically identical right-hand sides (v*v vs v^2, reordered terms, -(1/tau)*v vs -v/tau ...)."""
    if a == b:
        return True
    la, _, ra = a.partition("=")
    lb, _, rb = b.partition("=")
    if la != lb or not ra or not rb:
        return False
    try:
        import sympy
        diff = _rhs_expr(ra) - _rhs_expr(rb)
        return sympy.simplify(diff) == 0 or bool(diff.equals(0))
    except Exception:      # unparsable user text: not one of the compiled fields
        return False


def _normalise(eq_list: List[str]) -> List[str]:
    """The compiled vector field's canonical equation list that `eq_list` (canonical text, equation order = state order) is
    algebraically identical to, or `eq_list` itself.  Variable and parameter names must be the template's: they select ABI slots."""
    if eq_list in _KNOWN_FIELDS:
        return eq_list
    for known in _KNOWN_FIELDS:
        if len(known) == len(eq_list) and all(_same_equation(a, b) for a, b in zip(eq_list, known)):
            return known
    return eq_list


_BUILTIN_OPS: Dict[str, OperatorDef] = {
    "li_op": OperatorDef("li_op", ["v' = -v/tau + k*r_in + I_ext + eta"],
                         dict(v=("output", 0.0), tau=10.0, k=1.0, eta=0.0, r_in=("input", 0.0), I_ext=("input", 0.0))),
    "sigmoid_op": OperatorDef("sigmoid_op", ["r = r_max / (1 + exp(s*(v0-v)))"],
                              dict(r=("output", 0.0), r_max=1.0, s=1.0, v0=0.0, v=("input", 0.0))),
    "tanh_op": OperatorDef("tanh_op", ["r = tanh(v)"], dict(r=("output", 0.0), v=("input", 0.0))),
    "qif_op": OperatorDef("qif_op", ["v' = (v^2 + eta + I_ext)/tau + k*s_in", "s' = -s/tau_s + spike"],
                          dict(s=("output", 0.0), v=("variable", -2.0), tau=1.0, k=1.0, tau_s=1.0, eta=-5.0,
                               I_ext=("input", 0.0), spike=("input", 0.0), s_in=("input", 0.0))),
    "qif_sfa_op": OperatorDef("qif_sfa_op", ["v' = (v^2 + eta - x + I_ext)/tau + k*s_in", "s' = -s/tau_s + spike",
                                             "x' = -x/tau_x + alpha*spike"],
                              dict(s=("output", 0.0), v=("variable", -2.0), x=("variable", 0.0), tau=1.0, k=1.0, tau_s=1.0,
                                   eta=-5.0, alpha=1.0, tau_x=10.0, I_ext=("input", 0.0), spike=("input", 0.0),
                                   s_in=("input", 0.0))),
    "lif_op": OperatorDef("lif_op", ["v' = -v/tau + k*s_in + I_ext + eta", "s' = -s/tau_s + spike + s_ext"],
                          dict(s=("output", 0.0), v=("variable", 0.0), tau=10.0, k=1.0, eta=0.0, tau_s=0.5,
                               I_ext=("input", 0.0), spike=("input", 0.0), s_in=("input", 0.0), s_ext=("input", 0.0))),
}
_BUILTIN_OPS["ik_op"] = OperatorDef(
    "ik_op", ["v' = (k*(v-v_r)*(v-v_theta) - u + I_ext + eta + g*s_in*(E_r - v)) / C", "u' = (b*(v-v_r) - u) / tau_u + kappa*spike",
              "s' = -s/tau_s + spike"],
    dict(s=("output", 0.0), v=("variable", -60.0), u=("variable", 0.0), C=100.0, k=0.7, v_r=-60.0, v_theta=-40.0, eta=0.0, g=1.0,
         E_r=0.0, b=-2.0, tau_u=33.33, kappa=10.0, tau_s=6.0, I_ext=("input", 0.0), spike=("input", 0.0), s_in=("input", 0.0)))
# iku_op (ik.yaml:33-39): ik_op with  b*(v-v_r) -> b*(mean(v)-v_r)  and  kappa*spike -> kappa*mean(spike)
_BUILTIN_OPS["iku_op"] = OperatorDef(
    "iku_op", [_BUILTIN_OPS["ik_op"].equations[0], "u' = (b*(mean(v)-v_r) - u) / tau_u + kappa*mean(spike)", "s' = -s/tau_s + spike"],
    dict(_BUILTIN_OPS["ik_op"].variables))
# ik_biexp_op (ik.yaml:42-70): iku_op with a bi-exponential synapse  s' = -s/tau_d + x ,  x' = -x/tau_r + spike
_BUILTIN_OPS["ik_biexp_op"] = OperatorDef(
    "ik_biexp_op", [_BUILTIN_OPS["ik_op"].equations[0], _BUILTIN_OPS["iku_op"].equations[1], "s' = -s/tau_d + x", "x' = -x/tau_r + spike"],
    dict({k: v for k, v in _BUILTIN_OPS["ik_op"].variables.items() if k != "tau_s"}, x=("variable", 0.0), tau_r=2.0, tau_d=6.0))
_BUILTIN_NODES = {"tanh": ["li_op", "tanh_op"], "sigmoid": ["li_op", "sigmoid_op"], "qif": ["qif_op"],
                  "qif_sfa": ["qif_sfa_op"], "lif": ["lif_op"], "ik": ["ik_op"], "iku": ["iku_op"], "ik_biexp": ["ik_biexp_op"]}
_BUILTIN_MODULES = {"leaky_integrator": ["tanh", "sigmoid"], "qif": ["qif", "qif_sfa"], "lif": ["lif"], "ik": ["ik", "iku", "ik_biexp"]}

_SLOT = dict(tau=abi.RP_P_TAU, k=abi.RP_P_K, eta=abi.RP_P_ETA, tau_s=abi.RP_P_TAU_S, tau_x=abi.RP_P_TAU_X,
             alpha=abi.RP_P_ALPHA, r_max=abi.RP_P_RMAX, s=abi.RP_P_SIG_S, v0=abi.RP_P_V0,
             C=abi.RP_P_C, v_r=abi.RP_P_VR, v_theta=abi.RP_P_VTH, g=abi.RP_P_G, E_r=abi.RP_P_ER, b=abi.RP_P_B,
             tau_u=abi.RP_P_TAU_U, kappa=abi.RP_P_KAPPA)


def _val(v) -> float:
    return float(v[1]) if isinstance(v, tuple) else float(v)


def _spec_from_ops(ops: List[OperatorDef], force_jit: bool = False) -> TemplateSpec:
    """Recognise the operator combination by its equations and build the engine tables.  `force_jit`: skip the compiled fields
    (node semantics they do not implement, e.g. MultiSpikeResetNet)."""
    if force_jit:
        from . import jit
        return jit.unbound_spec(ops)
    eqs = [_normalise([_canon(e) for e in op.equations]) for op in ops]
    main = ops[0]
    mv = main.variables
    if len(ops) == 2 and eqs[0] == _LI and eqs[1] in (_TANH, _SIGMOID):
        act = ops[1]
        is_tanh = eqs[1] == _TANH
        params = {f"{main.name}/{p}": (_SLOT[p], _val(mv[p])) for p in ("tau", "k", "eta")}
        if not is_tanh:
            for p in ("r_max", "s", "v0"):
                params[f"{act.name}/{p}"] = (_SLOT[p], _val(act.variables[p]))
        return TemplateSpec(
            name="li_tanh" if is_tanh else "li_sigmoid", model=abi.RP_LI_TANH if is_tanh else abi.RP_LI_SIGMOID,
            ops=(main.name, act.name), state_vars=[(f"{main.name}/v", _val(mv["v"]))], params=params,
            source_var=f"{act.name}/r", target_var=f"{main.name}/r_in", input_vars={f"{main.name}/I_ext": 0},
            spike_var=None, out_vars={f"{main.name}/v": abi.RP_VAR_V, f"{act.name}/r": abi.RP_VAR_R})
    if len(ops) == 1 and eqs[0] in (_IK, _IKU):
        o = main.name
        pn = ["C", "k", "v_r", "v_theta", "eta", "g", "E_r", "b", "tau_u", "kappa", "tau_s"]
        mean_field = eqs[0] == _IKU
        return TemplateSpec(
            name="iku" if mean_field else "ik", model=abi.RP_IKU if mean_field else abi.RP_IK, ops=(o,), state_vars=[(f"{o}/{v}", _val(mv[v])) for v in ("v", "u", "s")],
            params={f"{o}/{p}": (_SLOT[p], _val(mv[p])) for p in pn},
            source_var=f"{o}/s", target_var=f"{o}/s_in", input_vars={f"{o}/I_ext": 0}, spike_var=f"{o}/spike",
            out_vars={f"{o}/v": abi.RP_VAR_V, f"{o}/s": abi.RP_VAR_S, f"{o}/u": abi.RP_VAR_X},
            planes={f"{o}/v": 0, f"{o}/s": 1, f"{o}/u": 2}, fold_param=f"{o}/g")
    if len(ops) == 1 and eqs[0] == _IK_BIEXP:
        # reference order of y: v, u, s, x (equation order); engine planes: v, s, u, x.  The decay / rise time constants travel in
        # the tau_s / tau_x slots of the ABI.
        o = main.name
        slot = dict(_SLOT, tau_d=abi.RP_P_TAU_S, tau_r=abi.RP_P_TAU_X)
        pn = ["C", "k", "v_r", "v_theta", "eta", "g", "E_r", "b", "tau_u", "kappa", "tau_d", "tau_r"]
        return TemplateSpec(
            name="ik_biexp", model=abi.RP_IK_BIEXP, ops=(o,), state_vars=[(f"{o}/{v}", _val(mv[v])) for v in ("v", "u", "s", "x")],
            params={f"{o}/{p}": (slot[p], _val(mv[p])) for p in pn},
            source_var=f"{o}/s", target_var=f"{o}/s_in", input_vars={f"{o}/I_ext": 0}, spike_var=f"{o}/spike",
            out_vars={f"{o}/v": abi.RP_VAR_V, f"{o}/s": abi.RP_VAR_S, f"{o}/u": abi.RP_VAR_X},
            planes={f"{o}/v": 0, f"{o}/s": 1, f"{o}/u": 2, f"{o}/x": 3}, fold_param=f"{o}/g")
    if len(ops) == 1 and eqs[0] in (_QIF, _QIF_SFA, _LIF):
        o = main.name
        if eqs[0] == _QIF:
            name, model, sv, pn = "qif", abi.RP_QIF, ["v", "s"], ["tau", "k", "eta", "tau_s"]
        elif eqs[0] == _QIF_SFA:
            name, model, sv, pn = "qif_sfa", abi.RP_QIF_SFA, ["v", "s", "x"], ["tau", "k", "eta", "tau_s", "tau_x", "alpha"]
        else:
            name, model, sv, pn = "lif", abi.RP_LIF, ["v", "s"], ["tau", "k", "eta", "tau_s"]
        inputs = {f"{o}/I_ext": 0}
        if name == "lif":
            inputs[f"{o}/s_ext"] = 1
        return TemplateSpec(
            name=name, model=model, ops=(o,), state_vars=[(f"{o}/{v}", _val(mv[v])) for v in sv],
            params={f"{o}/{p}": (_SLOT[p], _val(mv[p])) for p in pn},
            source_var=f"{o}/s", target_var=f"{o}/s_in", input_vars=inputs, spike_var=f"{o}/spike",
            out_vars={f"{o}/{v}": i for i, v in enumerate(sv)})
    # none of the compiled fields: generate the kernels at run time (rectipy_b200/jit.py), as PyRates does for the reference
    from . import jit
    return jit.unbound_spec(ops)


# ------------------------------------------------------------------------------------------------------------
# PyRates-style YAML parsing (subset: operator/node templates, `base` inheritance, `equations: replace/add`)
# ------------------------------------------------------------------------------------------------------------
_VAR_RE = re.compile(r"^\s*(input|output|variable)\s*(?:\(\s*([-+0-9.eE]+)?\s*\))?\s*$")


def _parse_var(v):
    if isinstance(v, (int, float)):
        return float(v)
    if isinstance(v, str):
        m = _VAR_RE.match(v)
        if m:
            return (m.group(1), float(m.group(2)) if m.group(2) else 0.0)
        return float(v)
    raise ValueError(f"cannot parse template variable definition {v!r}")


def _find_yaml(module_path: str) -> Optional[str]:
    """'pkg.sub.file' -> path of pkg/sub/file.yaml|yml if it exists (cwd, sys.path, or an importable package)."""
    import sys
    rel = module_path.replace(".", os.sep)
    for root in [os.getcwd()] + list(sys.path):
        for ext in (".yaml", ".yml"):
            p = os.path.join(root or ".", rel + ext)
            if os.path.isfile(p):
                return p
    top = module_path.split(".")[0]
    try:
        spec = importlib.util.find_spec(top)
    except (ImportError, ValueError):
        spec = None
    if spec and spec.submodule_search_locations:
        for loc in spec.submodule_search_locations:
            for ext in (".yaml", ".yml"):
                p = os.path.join(os.path.dirname(loc), rel + ext)
                if os.path.isfile(p):
                    return p
    return None


def _load_operator(path: str, name: str, docs: dict, seen=()) -> OperatorDef:
    """Resolve operator `name` (possibly 'pkg.file.op') with `base:` inheritance."""
    if name in seen:
        raise ValueError(f"circular template inheritance at {name}")
    if name not in docs:
        if "." in name:
            mod, leaf = name.rsplit(".", 1)
            return _resolve_operator_path(mod, leaf)
        if name in _BUILTIN_OPS:
            return _BUILTIN_OPS[name]
        raise KeyError(f"operator template {name} not found in {path}")
    body = docs[name]
    base = str(body.get("base", "OperatorTemplate"))
    if base.split(".")[-1] == "OperatorTemplate":
        eqs = body.get("equations", [])
        eqs = [eqs] if isinstance(eqs, str) else list(eqs)
        variables = {k: _parse_var(v) for k, v in (body.get("variables") or {}).items()}
        return OperatorDef(name, eqs, variables)
    parent = _load_operator(path, base, docs, seen + (name,))
    eqs = list(parent.equations)
    eq_mod = body.get("equations") or {}
    if isinstance(eq_mod, dict):
        for old, new in (eq_mod.get("replace") or {}).items():
            eqs = [e.replace(str(old), str(new)) for e in eqs]
        add = eq_mod.get("add") or []
        eqs += [add] if isinstance(add, str) else list(add)
    else:
        eqs = [eq_mod] if isinstance(eq_mod, str) else list(eq_mod)
    variables = dict(parent.variables)
    variables.update({k: _parse_var(v) for k, v in (body.get("variables") or {}).items()})
    return OperatorDef(name, eqs, variables)


def _resolve_operator_path(module_path: str, leaf: str) -> OperatorDef:
    path = _find_yaml(module_path)
    if path is None:
        if leaf in _BUILTIN_OPS:
            return _BUILTIN_OPS[leaf]
        raise FileNotFoundError(f"template file for {module_path} not found")
    import yaml
    with open(path) as fh:
        docs = yaml.safe_load(fh) or {}
    return _load_operator(path, leaf, docs)


def resolve_template(node, force_jit: bool = False) -> TemplateSpec:
    """`node`: dotted template path as accepted by the reference (e.g. "neuron_model_templates.spiking_neurons.qif.qif")."""
    if isinstance(node, TemplateSpec):
        return node
    if not isinstance(node, str):
        raise NotImplementedError("rectipy_b200 accepts template paths (str) or TemplateSpec objects; PyRates "
                                  f"NodeTemplate/CircuitTemplate instances are not supported (got {type(node).__name__}).")
    if "." not in node and "/" not in node:
        if node in _BUILTIN_NODES:
            return _spec_from_ops([_BUILTIN_OPS[o] for o in _BUILTIN_NODES[node]], force_jit)
        raise FileNotFoundError(f"Template {node} could not be found.")
    module_path, leaf = node.replace("/", ".").rsplit(".", 1)
    path = _find_yaml(module_path)
    if path is not None:
        import yaml
        with open(path) as fh:
            docs = yaml.safe_load(fh) or {}
        if leaf not in docs:
            raise AttributeError(f"Template {leaf} is not defined in {path}.")
        body = docs[leaf]
        op_names = body.get("operators") or []
        if isinstance(op_names, dict):
            op_names = list(op_names.keys())
        return _spec_from_ops([_load_operator(path, str(o), docs) for o in op_names], force_jit)
    mod = module_path.split(".")[-1]
    if module_path.split(".")[0] == "neuron_model_templates" and mod in _BUILTIN_MODULES:
        if leaf not in _BUILTIN_MODULES[mod]:
            raise AttributeError(f"Template {leaf} is not defined in {module_path}.")
        return _spec_from_ops([_BUILTIN_OPS[o] for o in _BUILTIN_NODES[leaf]], force_jit)
    raise FileNotFoundError(f"Template file {module_path} could not be found.")
