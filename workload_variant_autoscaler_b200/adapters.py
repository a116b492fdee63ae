"""Input packer and output path around the optimizer hot path.

The reference converts Kubernetes objects (ConfigMap strings, VariantAutoscaling CRs) into
``config.SystemData`` with the adapters of ``internal/utils/utils.go`` and turns the optimizer's
``AllocationSolution`` back into CR status + Prometheus gauges.  This module restates those adapters
on plain dicts shaped like the objects' JSON, so that a caller can go

    ConfigMaps + VariantAutoscalings --(this module)--> SystemSpec dict --Fleet.from_spec--> SoA fleet
        --Engine.solve--> winners --generate_solution / create_optimized_alloc--> CR status

Function names follow the reference (snake_case); each cites the lines it follows.  String -> float32
parsing is Go's ``strconv.ParseFloat(s, 32)``: correctly rounded straight to float32 (not through
float64), same accepted syntax, range errors reported.
"""
from __future__ import annotations

import os
import re
from fractions import Fraction

import numpy as np

__all__ = [
    "go_parse_float32", "check_value", "full_name", "create_system_data",
    "add_model_accelerator_profile_to_system_data", "add_server_info_to_system_data", "find_model_slo",
    "generate_solution", "create_optimized_alloc", "replica_metrics", "AdapterError",
]

ACCELERATOR_NAME_LABEL = "inference.optimization/acceleratorName"  # utils.go:296


class AdapterError(ValueError):
    """An error the reference returns from an adapter (``error`` results of utils.go)."""


# ----------------------------------------------------------------------------
# strconv.ParseFloat(s, 32)
# ----------------------------------------------------------------------------
_DEC = re.compile(r"([+-]?)(?:([0-9]+)(?:\.([0-9]*))?|\.([0-9]+))(?:[eE]([+-]?[0-9]+))?")
_HEX = re.compile(r"([+-]?)0[xX](?:([0-9a-fA-F]+)(?:\.([0-9a-fA-F]*))?|\.([0-9a-fA-F]+))[pP]([+-]?[0-9]+)")
_F32_MAX = Fraction(int(np.finfo(np.float32).max))
_F32_HALF_ULP_AT_MAX = Fraction(2) ** 103  # ulp(max) = 2^104
_BIG, _TINY = object(), object()  # magnitude far outside float32: decided from the exponent alone, in O(len)


def _underscore_ok(s: str) -> bool:
    """strconv's underscoreOK: '_' only between digits, or between a base prefix and a digit."""
    saw = "^"
    i = 1 if s[:1] in "+-" else 0
    hexa = False
    if len(s) - i >= 2 and s[i] == "0" and s[i + 1] in "bBoOxX":
        hexa = s[i + 1] in "xX"
        i += 2
        saw = "0"
    for c in s[i:]:
        if c in "0123456789" or (hexa and c in "abcdefABCDEF"):
            saw = "0"
        elif c == "_":
            if saw != "0":
                return False
            saw = "_"
        else:
            if saw == "_":
                return False
            saw = "!"
    return saw != "_"


def _clamp_exp(ex: str) -> int:
    """Exponent digits -> int without building huge integers (Go stops accumulating at 10000)."""
    neg = ex[:1] == "-"
    d = ex.lstrip("+-").lstrip("0")
    v = 10 ** 6 if len(d) > 6 else int(d or "0")
    return -v if neg else v


def _exact_value(s: str):
    """Exact rational value of a Go float literal; None if the syntax is not accepted; _BIG / _TINY (with the
    sign as second element of a tuple) when the exponent alone puts it out of float32's range."""
    if "_" in s:
        if not _underscore_ok(s):
            return None
        s = s.replace("_", "")
    m = _DEC.fullmatch(s)
    if m:
        sign, ip, fp, fp_only, ex = m.groups()
        frac = (fp or "") + (fp_only or "")
        digits = (ip or "") + frac
        if digits == "":
            return None
        e = _clamp_exp(ex or "0")
        sig = digits.lstrip("0")
        if sig:
            top = len(sig) - len(frac) + e  # value < 10^top, >= 10^(top-1)
            if top > 45:
                return (_BIG, sign)
            if top < -60:
                return (_TINY, sign)
        v = Fraction(int(digits), 10 ** len(frac)) * Fraction(10) ** e
        return -v if sign == "-" else v
    m = _HEX.fullmatch(s)
    if m:
        sign, ip, fp, fp_only, ex = m.groups()
        frac = (fp or "") + (fp_only or "")
        digits = (ip or "") + frac
        if digits == "":
            return None
        e = _clamp_exp(ex)
        sig = digits.lstrip("0")
        if sig:
            top = 4 * (len(sig) - len(frac)) + e  # value < 2^top, >= 2^(top-4)
            if top > 140:
                return (_BIG, sign)
            if top < -200:
                return (_TINY, sign)
        v = Fraction(int(digits, 16), 2 ** (4 * len(frac))) * Fraction(2) ** e
        return -v if sign == "-" else v
    return None


def go_parse_float32(s: str):
    """``strconv.ParseFloat(s, 32)`` -> ``(value: np.float32, err: str | None)``.

    err is ``"syntax"`` (value 0) or ``"range"`` (value +-Inf) as in Go's ``*NumError``.  The value is
    the float32 nearest to the literal (ties to even), computed exactly — rounding through float64
    first differs on literals that sit within 2^-29 relative of a float32 rounding boundary.
    """
    if not isinstance(s, str):
        return np.float32(0.0), "syntax"
    low = s.lower()
    body = low[1:] if low[:1] in "+-" else low
    if body in ("inf", "infinity"):
        return np.float32("-inf" if low.startswith("-") else "inf"), None
    if body == "nan" and low == "nan":  # Go accepts "nan" without a sign only (plus case variants)
        return np.float32("nan"), None
    if not s.isascii():
        return np.float32(0.0), "syntax"
    x = _exact_value(s)
    if x is None:
        return np.float32(0.0), "syntax"
    if isinstance(x, tuple):  # decided from the exponent: overflow is a range error, underflow rounds to +-0
        kind, sign = x
        if kind is _BIG:
            return np.float32("-inf" if sign == "-" else "inf"), "range"
        return np.float32(-0.0 if sign == "-" else 0.0), None
    if x == 0:
        return np.float32(-0.0 if s.startswith("-") else 0.0), None
    ax = abs(x)
    if ax >= _F32_MAX + _F32_HALF_ULP_AT_MAX:
        return np.float32("-inf" if x < 0 else "inf"), "range"
    # nearest float32: start from the float64-rounded guess and fix it up exactly
    try:
        guess = np.float32(float(ax))
    except OverflowError:
        guess = np.float32(np.finfo(np.float32).max)
    if not np.isfinite(guess):
        guess = np.float32(np.finfo(np.float32).max)
    with np.errstate(over="ignore"):
        cands = {float(guess), float(np.nextafter(guess, np.float32(0.0))), float(np.nextafter(guess, np.float32(np.inf)))}
    cands = [c for c in cands if np.isfinite(c)]

    def key(c):
        d = abs(Fraction(c) - ax)
        even = (np.float32(c).view(np.uint32) & np.uint32(1)) == 0
        return (d, 0 if even else 1)

    best = np.float32(min(cands, key=key))
    return (np.float32(-best) if x < 0 else best), None


def check_value(x: float) -> bool:
    """utils.go:341-343: valid = not NaN and not infinite."""
    return not (np.isnan(x) or np.isinf(x))


def full_name(name: str, namespace: str) -> str:
    """utils.go:334-336."""
    return f"{name}:{namespace}"


# ----------------------------------------------------------------------------
# ConfigMaps -> SystemData  (utils.go:108-182)
# ----------------------------------------------------------------------------
def _parse_service_class(text: str):
    import yaml

    doc = yaml.safe_load(text)
    if not isinstance(doc, dict):
        raise AdapterError("service class entry is not a mapping")
    data = doc.get("data") or []
    if not isinstance(data, list):
        raise AdapterError("service class data is not a list")
    entries = []
    for e in data:
        if not isinstance(e, dict):
            raise AdapterError("service class data entry is not a mapping")
        # interfaces.ServiceClassEntry: model string, slo-tpot int, slo-ttft int (types.go:20-24)
        tpot, ttft = e.get("slo-tpot", 0), e.get("slo-ttft", 0)
        if isinstance(tpot, bool) or isinstance(ttft, bool) or not isinstance(tpot, int) or not isinstance(ttft, int):
            raise AdapterError("slo-tpot / slo-ttft must be integers")
        entries.append({"model": str(e.get("model", "")), "slo-tpot": tpot, "slo-ttft": ttft})
    prio = doc.get("priority", 0)
    if isinstance(prio, bool) or not isinstance(prio, int):
        raise AdapterError("priority must be an integer")
    return {"name": str(doc.get("name", "")), "priority": prio, "data": entries}


def create_system_data(accelerator_cm: dict, service_class_cm: dict) -> dict:
    """``CreateSystemData`` (utils.go:108-182): ConfigMap contents -> ``{"spec": SystemSpec}``.

    accelerator_cm: name -> {"device": ..., "cost": "<float>"}; an unparsable cost skips the accelerator.
    service_class_cm: key -> YAML text of one service class; an unparsable entry is skipped.
    Unlimited mode, empty capacity, no models / servers yet.
    """
    accelerators = []
    for name, val in accelerator_cm.items():
        cost, err = go_parse_float32(val.get("cost", ""))
        if err is not None:
            continue  # "failed to parse accelerator cost in configmap, skipping accelerator"
        accelerators.append({"name": name, "type": val.get("device", ""), "multiplicity": 1, "power": {},
                             "cost": float(cost)})
    classes = []
    for _key, text in service_class_cm.items():
        try:
            sc = _parse_service_class(text)
        except Exception:  # noqa: BLE001 - "failed to parse service class data, skipping service class"
            continue
        classes.append({
            "name": sc["name"], "priority": sc["priority"],
            "modelTargets": [{"model": e["model"], "slo-itl": float(np.float32(e["slo-tpot"])),
                              "slo-ttft": float(np.float32(e["slo-ttft"]))} for e in sc["data"]],
        })
    return {"spec": {
        "acceleratorData": {"accelerators": accelerators},
        "modelData": {"models": []},
        "serviceClassData": {"serviceClasses": classes},
        "serverData": {"servers": []},
        "optimizerData": {"optimizer": {"unlimited": True}},
        "capacityData": {"count": []},
    }}


def find_model_slo(service_class_cm: dict, target_model: str):
    """``FindModelSLO`` (utils.go:369-383) -> (entry, class name); raises if a class fails to parse or the
    model is in no class."""
    for key, text in service_class_cm.items():
        try:
            sc = _parse_service_class(text)
        except Exception as exc:  # noqa: BLE001
            raise AdapterError(f"failed to parse {key}: {exc}") from exc
        for e in sc["data"]:
            if e["model"] == target_model:
                return e, sc["name"]
    raise AdapterError(f'model "{target_model}" not found in any service class')


# ----------------------------------------------------------------------------
# VariantAutoscaling -> SystemData  (utils.go:185-311)
# ----------------------------------------------------------------------------
def add_model_accelerator_profile_to_system_data(sd: dict, model_name: str, profile: dict) -> None:
    """``AddModelAcceleratorProfileToSystemData`` (utils.go:185-234).  profile: the CR's AcceleratorProfile
    (``acc``, ``accCount``, ``maxBatchSize``, ``perfParms.{decodeParms,prefillParms}`` string maps)."""
    perf = profile.get("perfParms", {}) or {}
    decode, prefill = perf.get("decodeParms", {}) or {}, perf.get("prefillParms", {}) or {}
    if len(decode) < 2:
        raise AdapterError("length of decodeParms should be 2")
    vals = {}
    for k in ("alpha", "beta"):
        v, err = go_parse_float32(decode.get(k, ""))
        if err is not None:
            raise AdapterError(f'strconv.ParseFloat: parsing "{decode.get(k, "")}": {err}')
        vals[k] = v
    if len(prefill) < 2:
        raise AdapterError("length of prefillParms should be 2")
    for k in ("gamma", "delta"):
        v, err = go_parse_float32(prefill.get(k, ""))
        if err is not None:
            raise AdapterError(f'strconv.ParseFloat: parsing "{prefill.get(k, "")}": {err}')
        vals[k] = v
    sd["spec"]["modelData"]["models"].append({
        "name": model_name, "acc": profile.get("acc", ""), "accCount": int(profile.get("accCount", 0)),
        "maxBatchSize": int(profile.get("maxBatchSize", 0)),
        "decodeParms": {"alpha": float(vals["alpha"]), "beta": float(vals["beta"])},
        "prefillParms": {"gamma": float(vals["gamma"]), "delta": float(vals["delta"])},
    })


def _load_value(s) -> np.float32:
    """ParseFloat(...,32) with the adapter's fallback: error or NaN/Inf -> 0 (utils.go:244-252)."""
    v, err = go_parse_float32(s if isinstance(s, str) else "")
    if err is not None or not check_value(float(v)):
        return np.float32(0.0)
    return v


def _go_int(x: np.float32) -> int:
    """Go's int(float64(x)) for finite x: truncation toward zero."""
    return int(np.trunc(float(x)))


def add_server_info_to_system_data(sd: dict, va: dict, class_name: str, *, environ=None) -> None:
    """``AddServerInfoToSystemData`` (utils.go:237-311).  va: the VariantAutoscaling object as a dict
    (``metadata.{name,namespace,labels}``, ``spec.{modelID,modelProfile.accelerators}``,
    ``status.currentAlloc`` with string-valued load / cost / latency fields)."""
    environ = os.environ if environ is None else environ
    meta, spec = va.get("metadata", {}) or {}, va.get("spec", {}) or {}
    cur = (va.get("status", {}) or {}).get("currentAlloc", {}) or {}
    load = cur.get("load", {}) or {}
    arrival = _load_value(load.get("arrivalRate"))
    out_tok = _load_value(load.get("avgOutputTokens"))
    in_tok = _load_value(load.get("avgInputTokens"))
    allocation = {
        "accelerator": cur.get("accelerator", ""), "numReplicas": int(cur.get("numReplicas", 0)),
        "maxBatch": int(cur.get("maxBatch", 0)),
        "cost": float(_load_value(cur.get("variantCost"))),
        "itlAverage": float(_load_value(cur.get("itlAverage"))),
        "ttftAverage": float(_load_value(cur.get("ttftAverage"))),
        "load": {"arrivalRate": float(arrival), "avgInTokens": _go_int(in_tok), "avgOutTokens": _go_int(out_tok)},
    }
    min_replicas = 0 if environ.get("WVA_SCALE_TO_ZERO") == "true" else 1  # scale to zero is off by default
    server = {
        "name": full_name(meta.get("name", ""), meta.get("namespace", "")), "class": class_name,
        "model": spec.get("modelID", ""), "keepAccelerator": True, "minNumReplicas": min_replicas,
        "currentAlloc": allocation, "desiredAlloc": {},
    }
    acc_name = (meta.get("labels", {}) or {}).get(ACCELERATOR_NAME_LABEL, "")
    max_batch = 0
    for ap in (spec.get("modelProfile", {}) or {}).get("accelerators", []) or []:
        if ap.get("acc", "") == acc_name:
            max_batch = int(ap.get("maxBatchSize", 0))
            break
    if max_batch > 0:
        server["maxBatchSize"] = max_batch
    sd["spec"]["serverData"]["servers"].append(server)


# ----------------------------------------------------------------------------
# Output path  (core/system.go:303-319, utils.go:314-331, internal/metrics/metrics.go:103-126)
# ----------------------------------------------------------------------------
def generate_solution(fleet, winners) -> dict:
    """``System.GenerateSolution`` (system.go:303-319) from the engine's winner block: server name ->
    AllocationData (+ the server's load); servers without an allocation are absent."""
    def col(name):  # Allocs (engine) or a structured array (oracle)
        return winners[name] if isinstance(winners, np.ndarray) else getattr(winners, name)

    feasible, acc, replicas, batch = col("feasible"), col("acc"), col("replicas"), col("batch")
    cost, itl, ttft = col("cost"), col("itl"), col("ttft")
    out = {}
    for i, name in enumerate(fleet.server_names):
        if not int(feasible[i]):
            continue
        a = int(acc[i])
        out[name] = {
            "accelerator": fleet.acc_names[a] if 0 <= a < len(fleet.acc_names) else "",
            "numReplicas": int(replicas[i]), "maxBatch": int(batch[i]),
            "cost": float(cost[i]), "itlAverage": float(itl[i]), "ttftAverage": float(ttft[i]),
            "load": {"arrivalRate": float(fleet.srv_arrival_rpm[i]), "avgInTokens": int(fleet.srv_in_tokens[i]),
                     "avgOutTokens": int(fleet.srv_out_tokens[i])},
        }
    return {"spec": out}


def create_optimized_alloc(name: str, namespace: str, allocation_solution: dict, *, now=None) -> dict:
    """``CreateOptimizedAlloc`` (utils.go:314-331): the CR status block for one variant."""
    import datetime

    server = full_name(name, namespace)
    data = allocation_solution.get("spec", {}).get(server)
    if data is None:
        raise AdapterError(f"server {server} not found")
    now = now or datetime.datetime.now(datetime.timezone.utc)
    return {"lastRunTime": now.isoformat(), "accelerator": data["accelerator"], "numReplicas": data["numReplicas"]}


def replica_metrics(current: int, desired: int) -> dict:
    """Gauge values of ``EmitReplicaMetrics`` (metrics.go:103-126): 0 -> N is reported as ratio N."""
    ratio = float(desired) if current == 0 else float(desired) / float(current)
    return {"current_replicas": float(current), "desired_replicas": float(desired), "desired_ratio": ratio}
