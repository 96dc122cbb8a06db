"""Shared helpers for the test-suite (oracle side)."""
import re

from oracle import oracle, tsa_config, alphabets

OPN = oracle.OP_INDEX
PSD = {"Reference": 0, "Query": 1, "Forward": 0, "Reverse": 1}


def config_from_dict(d):
    """Config built in code by a reference test (kats.json 'compute_cost')."""
    a = d["alphabet"]
    chars = alphabets.chars(a)
    A = len(chars)
    cfg = tsa_config.Config(alphabet=a, chars=chars)
    cfg.left_flank_length = d["left_flank_length"]
    cfg.right_flank_length = d["right_flank_length"]
    cfg.base = list(d["base"])
    cfg.fns = [[tuple(p) for p in f] for f in d["fns"]]
    cfg.tables = []
    for name, t in zip(tsa_config.TABLE_NAMES, d["tables"]):
        sub = [[t["match"] if x == y else t["sub"] for y in range(A)] for x in range(A)]
        cfg.tables.append(tsa_config.Table(name, sub, list(t["open"]), list(t["ext"])))
    return cfg


def ops_from_json(alignment):
    """[[count, name, {fields}] | [count, name]] -> [oracle.Op]"""
    out = []
    for item in alignment:
        count, name = item[0], item[1]
        f = item[2] if len(item) > 2 else {}
        if name == "TemplateSwitchEntrance":
            out.append(oracle.Op(count, OPN[name], PSD[f["primary"]], PSD[f["secondary"]], PSD[f["direction"]], f["first_offset"]))
        elif name == "TemplateSwitchExit":
            out.append(oracle.Op(count, OPN[name], 0, 0, 0, f["anti_primary_gap"]))
        else:
            out.append(oracle.Op(count, OPN[name]))
    return out


def ops_from_toml(alignment):
    """Parsed golden TOML `alignment` array -> [oracle.Op] (alignment_type.rs:11-75 serde layout)."""
    out = []
    for count, op in alignment:
        if isinstance(op, str):
            out.append(oracle.Op(count, OPN[op]))
        elif "TemplateSwitchEntrance" in op:
            f = op["TemplateSwitchEntrance"]
            out.append(oracle.Op(count, OPN["TemplateSwitchEntrance"], PSD[f["primary"]], PSD[f["secondary"]], PSD[f["direction"]], f["first_offset"]))
        else:
            out.append(oracle.Op(count, OPN["TemplateSwitchExit"], 0, 0, 0, op["TemplateSwitchExit"]["anti_primary_gap"]))
    return out


def clean_record(seq, skip="-"):
    """tsalign/src/align.rs:320-335: drop skip characters, then uppercase."""
    return "".join(c for c in seq if c not in skip).upper()


def split_embedded(seq):
    """tsalign/src/align.rs:338-379: '|' marks offset and limit; returns (clean sequence, offset, limit)."""
    parts = seq.split("|")
    if len(parts) != 3:
        raise ValueError("expected exactly two '|' characters")
    return "".join(parts), len(parts[0]), len(parts[0]) + len(parts[1])


def parse_rq_ranges(s):
    m = re.fullmatch(r"R(\d+)\.\.(\d+)Q(\d+)\.\.(\d+)", s.replace(" ", ""))
    return tuple(int(x) for x in m.groups())


def parse_config_any(text):
    """The shipped configs are written for dna-n (sample, ...) or dna (bench, experiments): pick the one that parses."""
    for alphabet in ("dna-n", "dna"):
        try:
            return tsa_config.parse(text, alphabet)
        except tsa_config.ConfigError:
            pass
    raise tsa_config.ConfigError("no alphabet fits")


def config_to_text(cfg):
    """oracle tsa_config.Config -> config.tsa text (the syntax of sample_tsa_config/config.tsa)."""
    def cost(v):
        return "inf" if v == tsa_config.INF else str(v)

    def idx(v):
        return "-inf" if v == tsa_config.I64_MIN else ("inf" if v == tsa_config.I64_MAX else str(v))

    out = ["# Limits", "", f"left_flank_length = {cfg.left_flank_length}", f"right_flank_length = {cfg.right_flank_length}", "", "# Base Cost", ""]
    for name, v in zip(tsa_config.BASE_NAMES, cfg.base):
        out.append(f"{name}_cost = {cost(v)}")
    out += ["", "# Jump Costs", ""]
    for name, pts in zip(tsa_config.FN_NAMES, cfg.fns):
        out += [name, " " + " ".join(idx(x) for x, _ in pts), " " + " ".join(cost(c) for _, c in pts), ""]
    chars = cfg.chars
    for t in cfg.tables:
        out += [f"# {t.name}", "", "SubstitutionCostTable", "  | " + " ".join(chars), "--+" + "-" * (2 * len(chars))]
        for r, ch in enumerate(chars):
            out.append(f"{ch} | " + " ".join(cost(v) for v in t.sub[r]))
        out += ["", "GapOpenCostVector", " " + " ".join(chars), " " + " ".join(cost(v) for v in t.open), "",
                "GapExtendCostVector", " " + " ".join(chars), " " + " ".join(cost(v) for v in t.ext), ""]
    return "\n".join(out) + "\n"
