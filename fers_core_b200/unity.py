"""Unity checks: host side.

Wire contract: ``analysis.unity_checks`` = list of ``UnityCheckDefinition``
(fers_core/types/pydantic_models.py:2202-2220) whose ``spec`` is ``{"Generic": GenericSpec}``
(2136-2151) or ``{"Ec3Steel": ...}`` (1448-1469); builders in fers_core/unity_checks.py:67-198.
Results: ``ResultsBundle.unity_check_results`` = list of ``UnityCheckResult`` (2005-2012) with
``EntityUnityResult`` rows (1472-1489) and status colours from ``thresholds`` [0.8, 0.95, 1.0].

This module compiles a Generic check into the tables ``csrc/unity.cu`` evaluates on the device
(variable descriptors + postfix programs + per-entity constants) and folds the per-batch envelopes.
Semantics where the closed engine is the only definition (stated, unpinned):
  * utilisation = demand / capacity per (entity, combination); the result of an entity is the
    combination with the largest utilisation (``governing_combination_id``);
  * member forces are the member's local section forces: Max/Min/MaxAbs from ``local_maximums`` /
    ``local_minimums`` (envelope over the member set for set entities), Start/End =
    ``local_start_forces`` / ``local_end_forces``, AtFraction = internal-force series at
    0, 1/4, 1/2, 3/4, 1 (nearest station);
  * MemberDeflection = displacement of the entity's nodes relative to the chord through its end nodes
    (example 100: equals the global mid-span deflection on rigid supports);
  * applicable combinations: ``load_combination_ids`` when given, else the combinations whose
    ``limit_state`` equals the check's (all combinations when either side has none).
  * ``of: ReferenceMember`` / ``ReferenceMember*`` node references follow ``Member.reference_member`` /
    ``reference_node`` of the entity's first member (example 101); an entity without one evaluates to NaN;
  * ``Displacement`` = largest absolute value of the component over the entity's nodes.
  * plate entities (``AllPlates`` / ``PlateElements`` / ``PlateSurfaces``): ``PlateStress`` = surface stress n/t +- 6m/t^2 at the element
    centroid, both surfaces, enveloped; QUAD4 elements (two TRI3 rows) envelope both halves.
Ec3Steel checks are compiled into the same postfix programs (ec3.py: EN 1993-1-1 6.2 cross-section rules, 6.3.1 flexural
buckling, 6.3.2 lateral-torsional buckling, linear 6.3.3 interaction).  Every entity result carries a trace with the value
of each variable (Ec3Steel: each rule under its clause label) at its governing combination.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

from . import _lib, ec3, wire

OPS = {"PUSHC": 0, "PUSHV": 1, "+": 2, "-": 3, "*": 4, "/": 5, "^": 6, "NEG": 7, "abs": 8, "sqrt": 9, "min": 10, "max": 11}
V_CONST, V_FORCE, V_EXPR, V_NODE_DISP, V_NODE_POS, V_DEFLECTION, V_FORCE_REF, V_ENTITY_DISP, V_PLATE = range(9)
PLATE_MEASURE = {"VonMises": 0, "PrincipalMax": 1, "PrincipalMin": 2}
NODE_REF = {"Start": 0, "End": 1, "ReferenceMemberStart": 2, "ReferenceMemberEnd": 3, "ReferenceNode": 4}
AGG = {"MaxAbs": 0, "Max": 1, "Min": 2, "Start": 3, "End": 4, "AtFraction": 5}
FORCE_COMP = {"N": 0, "Vy": 1, "Vz": 2, "Mx": 3, "My": 4, "Mz": 5}
DISP_COMP = {"Dx": 0, "Dy": 1, "Dz": 2, "Rx": 3, "Ry": 4, "Rz": 5, "Magnitude": 6}
AXIS = {"X": 0, "Y": 1, "Z": 2}
DEFL_COMP = {"LocalY": 0, "LocalZ": 1, "Magnitude": 2}
SECTION_FIELD = {"Area": "area", "Iy": "i_y", "Iz": "i_z", "J": "j", "Iw": "i_w", "WelY": "wel_y", "WelZ": "wel_z",
                 "WplY": "wpl_y", "WplZ": "wpl_z", "H": "h", "B": "b", "Asy": "a_sy", "Asz": "a_sz"}
MATERIAL_FIELD = {"E": "e_mod", "G": "g_mod", "Fy": "yield_stress", "Density": "density"}
MAX_VARS = 96
MAX_STACK = 16
STATUS = ("Green", "Yellow", "Orange", "Red")


class UnityError(ValueError):
    pass


# ---- expression compiler: infix -> postfix (op, arg) pairs -------------------------------------
_TOKEN = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)|([A-Za-z_][A-Za-z_0-9]*)|(.))")


def _tokens(text: str):
    out, pos = [], 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            break
        pos = m.end()
        if m.group(1) is not None:
            out.append(("num", float(m.group(1))))
        elif m.group(2) is not None:
            out.append(("id", m.group(2)))
        elif m.group(3).strip():
            out.append(("op", m.group(3)))
    return out


class _Parser:
    """expr := term (('+'|'-') term)* ; term := unary (('*'|'/') unary)* ; unary := '-' unary | power ;
    power := atom ('^' unary)? ; atom := number | name | func '(' args ')' | '(' expr ')'."""

    def __init__(self, text: str, var_index: dict, consts: list):
        self.t = _tokens(text)
        self.i = 0
        self.text = text
        self.var_index = var_index
        self.consts = consts
        self.code: list = []
        self.depth = 0
        self.max_depth = 0

    def _peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def _eat(self, kind=None, val=None):
        k, v = self._peek()
        if k is None or (kind and k != kind) or (val is not None and v != val):
            raise UnityError(f"expression {self.text!r}: unexpected {v!r}" if k else f"expression {self.text!r}: unexpected end")
        self.i += 1
        return v

    def _emit(self, op, arg=0, delta=0):
        self.code.append((OPS[op], arg))
        self.depth += delta
        self.max_depth = max(self.max_depth, self.depth)

    def parse(self):
        self.expr()
        if self.i != len(self.t):
            raise UnityError(f"expression {self.text!r}: trailing input at {self._peek()[1]!r}")
        if self.max_depth > MAX_STACK:
            raise UnityError(f"expression {self.text!r}: needs more than {MAX_STACK} stack slots")
        return self.code

    def expr(self):
        self.term()
        while self._peek() in (("op", "+"), ("op", "-")):
            op = self._eat()
            self.term()
            self._emit(op, 0, -1)

    def term(self):
        self.unary()
        while self._peek() in (("op", "*"), ("op", "/")):
            op = self._eat()
            self.unary()
            self._emit(op, 0, -1)

    def unary(self):
        if self._peek() == ("op", "-"):
            self._eat()
            self.unary()
            self._emit("NEG")
        elif self._peek() == ("op", "+"):
            self._eat()
            self.unary()
        else:
            self.power()

    def power(self):
        self.atom()
        if self._peek() == ("op", "^"):
            self._eat()
            self.unary()
            self._emit("^", 0, -1)

    def atom(self):
        k, v = self._peek()
        if k == "num":
            self._eat()
            self.consts.append(float(v))
            self._emit("PUSHC", len(self.consts) - 1, +1)
        elif k == "id":
            self._eat()
            if self._peek() == ("op", "("):
                self._eat()
                nargs = 1
                self.expr()
                while self._peek() == ("op", ","):
                    self._eat()
                    self.expr()
                    nargs += 1
                self._eat("op", ")")
                name = v.lower()
                if name in ("abs", "sqrt") and nargs == 1:
                    self._emit(name)
                elif name in ("min", "max") and nargs >= 2:
                    for _ in range(nargs - 1):
                        self._emit(name, 0, -1)
                elif name == "pow" and nargs == 2:
                    self._emit("^", 0, -1)
                else:
                    raise UnityError(f"expression {self.text!r}: unknown function {v}/{nargs}")
            else:
                if v not in self.var_index:
                    raise UnityError(f"expression {self.text!r}: unknown variable {v!r}")
                self._emit("PUSHV", self.var_index[v], +1)
        elif (k, v) == ("op", "("):
            self._eat()
            self.expr()
            self._eat("op", ")")
        else:
            raise UnityError(f"expression {self.text!r}: unexpected {v!r}")


def compile_expression(text: str, var_index: dict, consts: list) -> list:
    return _Parser(text, var_index, consts).parse()


# ---- check compilation -----------------------------------------------------------------------------
@dataclass
class CompiledCheck:
    definition: dict
    entities: list                     # EntityRef dicts
    ent_ptr: np.ndarray
    ent_members: np.ndarray
    ent_node_ptr: np.ndarray
    ent_nodes: np.ndarray
    frames: np.ndarray                 # [n_ent, 9]
    ent_ref: np.ndarray                # [n_ent, 4] reference member, its start / end node, reference node (-1 none)
    var_names: list
    var_desc: np.ndarray               # [n_vars, 4] int32
    code: np.ndarray                   # [n_code, 2] int32
    consts: np.ndarray
    ctab: np.ndarray                   # [n_ent, n_ct]
    demand: tuple                      # (offset, length)
    capacity: tuple
    combo_ids: set | None              # applicable combination ids (None = all)
    skipped: str | None = None         # reason when the check is not evaluated
    plates: bool = False               # entities are plate elements (ent_members = TRI3 rows)
    warnings: list = field(default_factory=list)
    spec: dict | None = None           # the Generic specification that was compiled (generated for Ec3Steel checks)
    labels: dict | None = None         # variable name -> trace label; None: every variable under its own name


def _chain(model, members: list):
    """Order the members of a set into a chain and return (members in chain order, node list)."""
    if len(members) == 1:
        m = members[0]
        return [m], [int(model.m_n1[m]), int(model.m_n2[m])]
    ends = {}
    for m in members:
        for n in (int(model.m_n1[m]), int(model.m_n2[m])):
            ends.setdefault(n, []).append(m)
    tips = sorted(n for n, ms in ends.items() if len(ms) == 1)
    if len(tips) != 2 or any(len(ms) > 2 for ms in ends.values()):
        nodes = []
        for m in members:                       # not a simple chain: keep wire order
            for n in (int(model.m_n1[m]), int(model.m_n2[m])):
                if n not in nodes:
                    nodes.append(n)
        return list(members), nodes
    first = int(model.m_n1[members[0]])
    start = first if first in tips else tips[0]
    order, nodes, cur, used = [], [start], start, set()
    while len(order) < len(members):
        nxt = [m for m in ends[cur] if m not in used]
        if not nxt:
            break
        m = nxt[0]
        used.add(m)
        order.append(m)
        cur = int(model.m_n2[m]) if int(model.m_n1[m]) == cur else int(model.m_n1[m])
        nodes.append(cur)
    if len(order) != len(members):
        return list(members), nodes
    return order, nodes


def _member_frame(model, m: int) -> np.ndarray:
    """Local axes of member m (rows x, y, z), rule of fers_core/members/member.py:309-354."""
    d = wire.member_axis(model, m)
    lx = d / np.linalg.norm(d)
    ref = np.array([0.0, 1.0, 0.0]) if abs(lx[1]) <= 1.0 - 1e-6 else np.array([0.0, 0.0, 1.0])
    lz = np.cross(lx, ref); lz /= np.linalg.norm(lz)
    ly = np.cross(lz, lx); ly /= np.linalg.norm(ly)
    roll = float(model.m_roll[m])
    if roll != 0.0:
        c, s = np.cos(roll), np.sin(roll)
        ly, lz = c * ly + s * lz, -s * ly + c * lz
    return np.concatenate([lx, ly, lz])


def _select_entities(model, sel: dict):
    kind = sel.get("type", "AllMembers")
    m_index = {int(i): k for k, i in enumerate(model.m_ids)}
    if kind == "AllMembers":
        return [({"type": "Member", "member_id": int(model.m_ids[k])}, [k]) for k in range(model.n_members)]
    if kind == "Members":
        out = []
        for i in sel.get("ids") or []:
            if int(i) not in m_index:
                raise UnityError(f"unity check: unknown member {i}")
            out.append(({"type": "Member", "member_id": int(i)}, [m_index[int(i)]]))
        return out
    if kind == "MemberSets":
        out = []
        for i in sel.get("ids") or []:
            if int(i) not in model.member_sets:
                raise UnityError(f"unity check: unknown member set {i}")
            out.append(({"type": "MemberSet", "memberset_id": int(i)}, list(model.member_sets[int(i)])))
        return out
    if kind == "MemberClassification":
        want = sel.get("value")
        return [({"type": "Member", "member_id": int(model.m_ids[k])}, [k]) for k in range(model.n_members)
                if model.m_classification[k] == want]
    if kind in ("AllPlates", "PlateElements"):
        rows_of = {}
        for row, pid in enumerate(model.p_ids):
            rows_of.setdefault(int(pid), []).append(row)
        ids = list(rows_of) if kind == "AllPlates" else [int(i) for i in sel.get("ids") or []]
        for i in ids:
            if i not in rows_of:
                raise UnityError(f"unity check: unknown plate element {i}")
        return [({"type": "PlateElement", "plate_element_id": i}, rows_of[i]) for i in ids]
    if kind == "PlateSurfaces":
        # EntitySelector6 / EntityRef4 (pydantic_models.py:314-319, 395-400): one entity per surface, made of the elements
        # meshed from it (PlateElement.source_surface_id); the envelope runs over all of them
        out = []
        for i in sel.get("ids") or []:
            rows = [int(r) for r in np.nonzero(np.asarray(model.p_surface) == int(i))[0]]
            if not rows:
                raise UnityError(f"unity check: plate surface {i} has no elements")
            out.append(({"type": "PlateSurface", "plate_surface_id": int(i)}, rows))
        return out
    raise UnityError(f"unity check: entity selector {kind!r} is not supported")


def compile_check(model, definition: dict) -> CompiledCheck:
    spec = definition.get("spec") or {}
    empty = dict(entities=[], ent_ptr=np.zeros(1, np.int32), ent_members=np.zeros(0, np.int32),
                 ent_node_ptr=np.zeros(1, np.int32), ent_nodes=np.zeros(0, np.int32), frames=np.zeros((0, 9)),
                 ent_ref=np.zeros((0, 4), np.int32),
                 var_names=[], var_desc=np.zeros((0, 4), np.int32), code=np.zeros((0, 2), np.int32), consts=np.zeros(1),
                 ctab=np.zeros((0, 1)), demand=(0, 0), capacity=(0, 0), combo_ids=None)
    if "Generic" not in spec and "Ec3Steel" not in spec:
        kind = next(iter(spec), "?")
        return CompiledCheck(definition=definition, skipped=f"{kind} checks are not evaluated by this engine", **empty)
    try:
        ents = _select_entities(model, definition.get("applies_to") or {"type": "AllMembers"})
    except UnityError as e:
        return CompiledCheck(definition=definition, skipped=str(e), **empty)
    ents = [(ref, ms) for ref, ms in ents if ms]
    n_ent = len(ents)
    plates = bool(ents) and ents[0][0]["type"] in ("PlateElement", "PlateSurface")
    labels = None
    entity_consts: dict = {}
    ec3_notes: list = []
    if "Ec3Steel" in spec:
        # EN 1993-1-1 member check: the rules become a Generic program over per-entity section constants (ec3.py)
        if plates:
            return CompiledCheck(definition=definition, skipped="Ec3Steel checks apply to members, not plates", **empty)
        ents = [(ref, ms) for ref, ms in ents if model.member_raw(ms[0])[0]]          # rigid links carry no section
        n_ent = len(ents)
        per = [ec3.entity_constants(model, spec["Ec3Steel"] or {}, _chain(model, ms)[0]) for _, ms in ents]
        entity_consts = {n: np.array([c[n] for c, _ in per], np.float64) for n in ec3.CONST_NAMES}
        ec3_notes = sorted({note for _, notes in per for note in notes})
        g = ec3.program(ec3.CONST_NAMES)
        labels = ec3.LABELS
    else:
        g = spec["Generic"]
    if plates:      # "members" = TRI3 rows of the plate, nodes = its corner nodes
        chains = [(ms, list(dict.fromkeys(int(n) for r in ms for n in model.p_nodes[r]))) for _, ms in ents]
    else:
        chains = [_chain(model, ms) for _, ms in ents]
    ent_ptr = np.zeros(n_ent + 1, np.int32)
    ent_node_ptr = np.zeros(n_ent + 1, np.int32)
    for i, (ms, ns) in enumerate(chains):
        ent_ptr[i + 1] = ent_ptr[i] + len(ms)
        ent_node_ptr[i + 1] = ent_node_ptr[i] + len(ns)
    ent_members = np.array([m for ms, _ in chains for m in ms], np.int32)
    ent_nodes = np.array([n for _, ns in chains for n in ns], np.int32)
    frames = (np.zeros((n_ent, 9)) if plates else
              np.array([_member_frame(model, ms[0]) for ms, _ in chains]).reshape(n_ent, 9))
    ent_ref = np.full((n_ent, 4), -1, np.int32)
    for i, (ms, _) in enumerate(chains if not plates else []):
        rm = int(model.m_ref_member[ms[0]]) if len(model.m_ref_member) else -1
        rn = int(model.m_ref_node[ms[0]]) if len(model.m_ref_node) else -1
        if rm >= 0:
            ent_ref[i, :3] = (rm, model.m_n1[rm], model.m_n2[rm])
        ent_ref[i, 3] = rn

    variables = g.get("variables") or []
    if len(variables) > MAX_VARS:
        raise UnityError(f"unity check {definition.get('id')}: more than {MAX_VARS} variables")
    names = [v["name"] for v in variables]
    var_index = {n: i for i, n in enumerate(names)}      # expressions may refer to any bound name
    order = _evaluation_order(variables, var_index)
    remap = {old: new for new, old in enumerate(order)}
    var_index = {n: remap[i] for n, i in var_index.items()}
    consts: list = [0.0]
    code: list = []
    ct_cols: list = []
    desc = np.zeros((len(variables), 4), np.int32)
    warnings: list = []
    multi = any(len(ms) > 1 for ms, _ in chains)
    for new, old in enumerate(order):
        src = variables[old]["source"]
        if "Expression" in src:
            prog = compile_expression(src["Expression"], var_index, consts)
            desc[new] = (V_EXPR, len(code), len(prog), 0)
            code += prog
            continue
        q = src.get("Quantity") or {}
        (qk, qv), = q.items() if len(q) == 1 else [("?", None)]
        if plates and qk not in ("Constant", "DimensionedConstant", "PlateStress", "Displacement"):
            return CompiledCheck(definition=definition, skipped=f"quantity {qk!r} is not defined for plate entities", **empty)
        if not plates and qk == "PlateStress":
            return CompiledCheck(definition=definition, skipped="PlateStress needs a plate selector (AllPlates / PlateElements / PlateSurfaces)",
                                 **empty)
        of_ref = isinstance(qv, dict) and qv.get("of") not in (None, "SelfEntity")
        if of_ref and qv.get("of") != "ReferenceMember":
            raise UnityError(f"unity check {definition.get('id')}: unknown target {qv.get('of')!r}")

        def target(i):      # member whose properties a constant quantity reads for entity i (-1: no reference member)
            return int(ent_ref[i, 0]) if of_ref else chains[i][0][0]
        if qk == "_EntityConstant":          # bound by the Ec3Steel program: one value per entity
            ct_cols.append(entity_consts[qv] if n_ent else np.zeros(0))
            desc[new] = (V_CONST, len(ct_cols) - 1, 0, 0)
        elif qk == "Constant":
            ct_cols.append(np.full(n_ent, float(qv)))
            desc[new] = (V_CONST, len(ct_cols) - 1, 0, 0)
        elif qk == "DimensionedConstant":
            ct_cols.append(np.full(n_ent, float(qv["value"])))
            desc[new] = (V_CONST, len(ct_cols) - 1, 0, 0)
        elif qk in ("Section", "Material"):
            table = SECTION_FIELD if qk == "Section" else MATERIAL_FIELD
            fld = table.get(qv.get("property"))
            if fld is None:
                raise UnityError(f"unity check {definition.get('id')}: unknown {qk} property {qv.get('property')!r}")
            col = np.full(n_ent, np.nan)
            for i in range(n_ent):
                raw = model.member_raw(target(i))[0 if qk == "Section" else 1] if target(i) >= 0 else None
                val = raw.get(fld) if raw else None
                col[i] = np.nan if val is None else float(val)
            ct_cols.append(col)
            desc[new] = (V_CONST, len(ct_cols) - 1, 0, 0)
        elif qk == "Geometry":
            col = np.zeros(n_ent)
            for i, (ms, _) in enumerate(chains):
                group = ([target(i)] if target(i) >= 0 else []) if of_ref else ms
                col[i] = sum(float(np.linalg.norm(wire.member_axis(model, m))) for m in group) \
                    if group else np.nan
            ct_cols.append(col)
            desc[new] = (V_CONST, len(ct_cols) - 1, 0, 0)
        elif qk == "MemberForce":
            comp = qv.get("component")
            if comp == "Bimoment":
                warnings.append("Bimoment is zero: the warping degree of freedom is not modelled")
                ct_cols.append(np.zeros(n_ent))
                desc[new] = (V_CONST, len(ct_cols) - 1, 0, 0)
                continue
            if comp not in FORCE_COMP:
                raise UnityError(f"unity check {definition.get('id')}: unknown force component {comp!r}")
            agg = qv.get("aggregation") or {"type": "MaxAbs"}
            at = AGG.get(agg.get("type"))
            if at is None:
                raise UnityError(f"unity check {definition.get('id')}: unknown aggregation {agg!r}")
            station = 0
            if at == AGG["AtFraction"]:
                if multi and not of_ref:
                    if labels is not None:       # Ec3Steel on a member set: quarter points of the FIRST member (noted)
                        if "C1 of a member-set entity uses the quarter points of its first member" not in warnings:
                            warnings.append("C1 of a member-set entity uses the quarter points of its first member")
                    else:
                        return CompiledCheck(definition=definition, **empty,
                                             skipped="AtFraction member forces on member-set entities are not evaluated")
                frac = float(agg.get("fraction", 0.5))
                station = int(round(frac * 4))
                if abs(station / 4 - frac) > 1e-12:
                    warnings.append(f"AtFraction {frac}: nearest of the stations 0, 1/4, 1/2, 3/4, 1 used")
            desc[new] = (V_FORCE_REF if of_ref else V_FORCE, FORCE_COMP[comp], at, station)
        elif qk in ("NodeDisplacement", "NodePosition"):
            at = qv.get("at")
            if at not in NODE_REF:
                raise UnityError(f"unity check {definition.get('id')}: unknown node reference {at!r}")
            if qk == "NodeDisplacement":
                desc[new] = (V_NODE_DISP, NODE_REF[at], DISP_COMP[qv.get("component")], 0)
            else:
                desc[new] = (V_NODE_POS, NODE_REF[at], AXIS[qv.get("axis")], 0)
        elif qk == "Displacement":
            desc[new] = (V_ENTITY_DISP, DISP_COMP[qv.get("component", "Magnitude")], 0, 0)
        elif qk == "PlateStress":
            if qv.get("measure", "VonMises") not in PLATE_MEASURE:
                raise UnityError(f"unity check {definition.get('id')}: unknown plate stress measure {qv.get('measure')!r}")
            desc[new] = (V_PLATE, PLATE_MEASURE[qv.get("measure", "VonMises")], 0, 0)
        elif qk == "MemberDeflection":
            agg = (qv.get("aggregation") or {"type": "MaxAbs"}).get("type")
            if agg == "AtFraction":
                return CompiledCheck(definition=definition, skipped="MemberDeflection AtFraction is not evaluated", **empty)
            if not multi:
                warnings.append("MemberDeflection is sampled at the nodes of the entity: a single member has no interior node")
            desc[new] = (V_DEFLECTION, DEFL_COMP[qv.get("component", "Magnitude")], AGG[agg], 0)
        else:
            return CompiledCheck(definition=definition, skipped=f"quantity {qk!r} is not evaluated by this engine", **empty)
    d_prog = compile_expression(g["demand"], var_index, consts)
    demand = (len(code), len(d_prog)); code += d_prog
    c_prog = compile_expression(g["capacity"], var_index, consts)
    capacity = (len(code), len(c_prog)); code += c_prog
    ctab = np.stack(ct_cols, axis=1) if ct_cols else np.zeros((n_ent, 1))

    ids = [int(i["id"] if isinstance(i, dict) else i) for i in (definition.get("load_combination_ids") or [])]
    combo_ids = set(ids) if ids else None
    if combo_ids is None and definition.get("limit_state") is not None:
        ls = [model.combo_limit_state[i] for i in range(len(model.combo_ids))]
        if any(x is not None for x in ls):
            combo_ids = {int(model.combo_ids[i]) for i, x in enumerate(ls) if x is None or x == definition["limit_state"]}
    return CompiledCheck(
        definition=definition, entities=[ref for ref, _ in ents], ent_ptr=ent_ptr, ent_members=ent_members,
        ent_node_ptr=ent_node_ptr, ent_nodes=ent_nodes, frames=frames, ent_ref=ent_ref, plates=plates,
        var_names=[names[o] for o in order], var_desc=desc,
        code=np.array(code, np.int32).reshape(-1, 2), consts=np.array(consts, np.float64), ctab=np.ascontiguousarray(ctab),
        demand=demand, capacity=capacity, combo_ids=combo_ids, warnings=warnings + ec3_notes, spec=g, labels=labels)


def _evaluation_order(variables: list, var_index: dict) -> list:
    """Topological order: an expression variable is evaluated after the variables it names."""
    deps = []
    for v in variables:
        src = v["source"]
        if "Expression" in src:
            deps.append({var_index[t[1]] for t in _tokens(src["Expression"]) if t[0] == "id" and t[1] in var_index})
        else:
            deps.append(set())
    order, done = [], set()
    while len(order) < len(variables):
        ready = [i for i in range(len(variables)) if i not in done and deps[i] <= done]
        if not ready:
            raise UnityError("unity check: cyclic variable definitions")
        for i in ready:
            order.append(i); done.add(i)
    return order


# ---- evaluation ---------------------------------------------------------------------------------------
class UnityAccumulator:
    """Runs every compiled check on each batch of result columns while they are on the device and
    keeps the running envelope per entity."""

    def __init__(self, model, plan, device):
        import torch
        self.torch = torch
        self.model, self.plan, self.dev = model, plan, device
        self.lib = _lib.load()
        self.checks = [compile_check(model, d) for d in (model.unity_checks or [])]
        # running envelope [utilisation, demand, capacity, plan column] and the variable values behind it, per entity, on the
        # device: folded there batch by batch (331 k members x 64 combinations: the host-side fold was 30 ms per batch)
        self._best_d = [torch.tensor([-np.inf, 0.0, 0.0, -1.0], dtype=torch.float64, device=device).repeat(len(c.entities), 1)
                        for c in self.checks]
        self._trace_d = [torch.zeros((len(c.entities), max(len(c.var_names), 1)), dtype=torch.float64, device=device)
                         for c in self.checks]
        self._dev = {}

    @property
    def best(self) -> list:
        return [b.cpu().numpy() for b in self._best_d]

    @property
    def trace(self) -> list:
        return [t.cpu().numpy() for t in self._trace_d]

    def reset(self) -> None:
        for b, t in zip(self._best_d, self._trace_d):
            b[:, 0] = -np.inf; b[:, 1:3] = 0.0; b[:, 3] = -1.0
            t.zero_()

    def __bool__(self):
        return any(c.skipped is None and len(c.entities) for c in self.checks)

    def _tensors(self, ci: int):
        if ci not in self._dev:
            c = self.checks[ci]
            t = lambda a: self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
            self._dev[ci] = dict(ent_ptr=t(c.ent_ptr), ent_members=t(c.ent_members if len(c.ent_members) else np.zeros(1, np.int32)),
                                 ent_node_ptr=t(c.ent_node_ptr), ent_nodes=t(c.ent_nodes if len(c.ent_nodes) else np.zeros(1, np.int32)),
                                 var_desc=t(c.var_desc if len(c.var_desc) else np.zeros((1, 4), np.int32)),
                                 code=t(c.code if len(c.code) else np.zeros((1, 2), np.int32)), consts=t(c.consts),
                                 ctab=t(c.ctab), frames=t(c.frames),
                                 ent_ref=t(c.ent_ref if len(c.ent_ref) else np.full((1, 4), -1, np.int32)))
        return self._dev[ci]

    def _p_data(self):
        if self.model.n_plates == 0:
            return None
        if "p_data" not in self._dev:
            self._dev["p_data"] = self.torch.from_numpy(np.ascontiguousarray(self.model.p_data, np.float64)).to(self.dev)
        return self._dev["p_data"]

    def accumulate(self, cols: list, d_xyz, d_u, rec: dict, stream: int) -> None:
        """cols: plan column of each of the first len(cols) device columns of this batch."""
        torch = self.torch
        nr = d_u.shape[1]
        for ci, c in enumerate(self.checks):
            if c.skipped is not None or not len(c.entities):
                continue
            mask = np.zeros(nr, np.uint8)
            for j, col in enumerate(cols):
                if self.plan.kinds[col] == "Loadcombination" and (c.combo_ids is None or int(self.plan.ids[col]) in c.combo_ids):
                    mask[j] = 1
            if not mask.any():
                continue
            if c.plates and "p_res" not in rec:
                raise UnityError("plate unity checks need the plate recovery")
            if not c.plates:
                for key in ("f_local", "sec", "lmax", "lmin"):
                    if key not in rec:
                        raise UnityError("unity checks need the detailed member recovery (detail=True)")
            T = self._tensors(ci)
            d_mask = torch.from_numpy(mask).to(self.dev)
            out = torch.empty((len(c.entities), 4), dtype=torch.float64, device=self.dev)
            vars_out = torch.zeros((len(c.entities), max(len(c.var_names), 1)), dtype=torch.float64, device=self.dev)
            P = _lib.ptr
            _lib.check(self.lib.fers_b200_unity_generic(
                len(c.entities), nr, len(c.var_desc), c.ctab.shape[1], P(d_mask), P(T["ent_ptr"]), P(T["ent_members"]),
                P(T["ent_node_ptr"]), P(T["ent_nodes"]), P(T["var_desc"]), P(T["code"]), P(T["consts"]), P(T["ctab"]),
                c.demand[0], c.demand[1], c.capacity[0], c.capacity[1], P(d_xyz), P(d_u), P(rec.get("f_local")),
                P(rec.get("sec")), P(rec.get("lmax")), P(rec.get("lmin")), P(T["frames"]), P(T["ent_ref"]), P(rec.get("p_res")),
                P(self._p_data()), P(out), P(vars_out), stream),
                "unity_generic")
            # fold_envelope on the device: batch column -> plan column, then best <- got where got is better
            cols_d = torch.tensor(cols, dtype=torch.float64, device=self.dev)
            have = out[:, 3] >= 0
            out[:, 3] = torch.where(have, cols_d[out[:, 3].clamp(min=0).long()], out[:, 3])
            best = self._best_d[ci]
            bu, gu = best[:, 0], out[:, 0]
            better = (torch.isnan(bu) & ~torch.isnan(gu)) | (~torch.isnan(gu) & ~torch.isnan(bu)
                                                           & ((gu > bu) | ((gu == bu) & (out[:, 3] < best[:, 3]))))
            take = have & ((best[:, 3] < 0) | better)
            self._best_d[ci] = torch.where(take[:, None], out, best)
            self._trace_d[ci] = torch.where(take[:, None], vars_out, self._trace_d[ci])

    def results(self) -> list:
        return [check_result(self.model, self.plan, c, b, t) for c, b, t in zip(self.checks, self.best, self.trace)]


def fold_envelope(best: np.ndarray, got: np.ndarray) -> np.ndarray:
    """best[e] <- got[e] where got has a column and a larger utilisation (ties: lower plan column); returns the mask
    of the entities that were replaced."""
    have = got[:, 3] >= 0
    none = best[:, 3] < 0
    bu, gu = best[:, 0], got[:, 0]
    better = (np.isnan(bu) & ~np.isnan(gu)) | (~np.isnan(gu) & ~np.isnan(bu) & ((gu > bu) | ((gu == bu) & (got[:, 3] < best[:, 3]))))
    take = have & (none | better)
    best[take] = got[take]
    return take


def _status(u: float, thresholds) -> str:
    t = list(thresholds or [0.8, 0.95, 1.0])
    if not np.isfinite(u):
        return "Red"
    for k, lim in enumerate(t[:3]):
        if u <= lim:
            return STATUS[k]
    return "Red"


def _trace(c: CompiledCheck, values, u, dem, cap) -> list:
    """CalcStep rows (pydantic_models.py:146-153): one per variable at the governing combination, then the verdict."""
    g = c.spec if c.spec is not None else c.definition["spec"]["Generic"]
    src = {v["name"]: v["source"] for v in g.get("variables") or []}
    if c.labels is not None:
        # generated program (Ec3Steel): the rule utilisations and a few intermediate values under their code-clause labels
        steps = [{"label": c.labels[n], "formula": str(src[n].get("Expression", "")), "substituted": "", "unit": "", "value": float(x)}
                 for n, x in zip(c.var_names, values) if n in c.labels and n != "gov"]
        steps.append({"label": "Governing", "formula": "max of the rules above", "substituted": "", "unit": "", "value": float(u)})
        return steps

    def formula(name):
        s_ = src.get(name, {})
        if "Expression" in s_:
            return str(s_["Expression"])
        q = s_.get("Quantity") or {}
        return next(iter(q), "")
    steps = [{"label": n, "formula": formula(n), "substituted": "", "unit": "", "value": float(x)}
             for n, x in zip(c.var_names, values)]
    steps.append({"label": "Demand", "formula": str(g["demand"]), "substituted": "", "unit": "", "value": float(dem)})
    steps.append({"label": "Capacity", "formula": str(g["capacity"]), "substituted": "", "unit": "", "value": float(cap)})
    steps.append({"label": "Utilization", "formula": "Demand / Capacity", "substituted": "", "unit": "", "value": float(u)})
    return steps


def check_result(model, plan, c: CompiledCheck, best: np.ndarray, trace: np.ndarray | None = None) -> dict:
    d = c.definition
    base = {"check_id": str(d.get("id")), "name": str(d.get("name", d.get("id"))), "limit_state": d.get("limit_state")}
    if c.skipped is not None or not len(c.entities):
        msg = c.skipped or "no entity matches the selector"
        return {**base, "max_utilization": 0.0, "status": "Green", "governing": None, "per_entity": [],
                "message": msg}
    per = []
    for i, ref in enumerate(c.entities):
        u, dem, cap, col = best[i]
        if col < 0:
            per.append({"entity": ref, "demand": 0.0, "capacity": None, "utilization": 0.0, "status": "Green",
                        "governing_combination_id": None, "location_x_frac": None,
                        "message": "no applicable load combination", "trace": []})
            continue
        uu = float(u) if np.isfinite(u) else float("inf")
        per.append({"entity": ref, "demand": float(dem), "capacity": float(cap), "utilization": uu,
                    "status": _status(u, d.get("thresholds")), "governing_combination_id": int(plan.ids[int(col)]),
                    "location_x_frac": None, "message": "; ".join(c.warnings) or None,
                    "trace": _trace(c, trace[i], uu, dem, cap) if trace is not None else []})
    gov = max(per, key=lambda r: r["utilization"])
    return {**base, "max_utilization": gov["utilization"], "status": gov["status"], "governing": gov, "per_entity": per}
