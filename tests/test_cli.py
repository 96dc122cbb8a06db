"""`tsalign-b200 align`: the reference's CLI surface (tsalign/src/align.rs:57-432) on the emulator build.
The five invocations of the reference's own integration tests (tsalign-tests/tests/integration.rs:6-29, which
assert exit status only) plus checks of the stdout block, the TOML layout and the rescoring of the written alignment."""
import os
import subprocess
import tomllib

import pytest

from oracle import oracle, tsa_config
from helpers import ops_from_toml
from template_switch_aligner_b200 import workloads
from emul_lib import emul

HERE = os.path.dirname(os.path.abspath(__file__))
CLI = os.path.join(HERE, "emul", "_build", "tsalign-b200-emul")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory, pairs, kats):
    emul()  # builds the emulator library and CLI
    d = tmp_path_factory.mktemp("cli")
    os.makedirs(d / "sample_tsa_config")
    (d / "sample_tsa_config" / "config.tsa").write_text(workloads.sample_config_text())
    os.makedirs(d / "test_files")
    for name in ("twin_a.fa", "twin_100_0.01.fa", "twin_embedded.fa", "twin_show_ts_indel1.fa", "twin_10_ts.fa"):
        (d / "test_files" / name).write_text(pairs[name]["raw"])
    (d / "test_files" / "reference_a.fa").write_text(kats["cli_smoke"]["reference_a.fa"])
    (d / "test_files" / "query_a.fa").write_text(kats["cli_smoke"]["query_a.fa"])
    return d


def run(workdir, *args):
    return subprocess.run([CLI, *args], cwd=workdir, capture_output=True, text=True, timeout=600)


def test_reference_integration_lines(workdir, kats):
    for line in kats["cli_smoke"]["lines"]:
        if "test_files/twin_100_0.01.fa" in line and "--memory-limit" in line:
            # --memory-limit bounds the HBM chunk here instead of aborting the search: still exit status 0
            pass
        r = run(workdir, *line)
        assert r.returncode == 0, (line, r.stderr)
        assert "Duration:" in r.stdout


def test_stdout_block_and_toml(workdir, toml_golden):
    r = run(workdir, "align", "-p", "test_files/twin_show_ts_indel1.fa", "-o", "out.toml")
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "CIGAR: 19=[TSQRR:[0,0]:[0,0]:-4:10=:10]15="   # equal-cost ranges as compute_ts_equal_cost_ranges reports them
    assert lines[1] == "Reached target with cost 2"
    assert lines[2:4] == ["Reference offset: 0", "Query offset: 0"]
    assert lines[4].startswith("Cost per base: 0.05") and lines[-1].startswith("Duration: ")
    doc = tomllib.loads((workdir / "out.toml").read_text())
    golden = toml_golden["twin_ari_chrX_146823507_146823598.toml"]["parsed"]
    assert set(doc) == set(golden) and set(doc["sequences"]) == set(golden["sequences"]) and set(doc["result"]) == set(golden["result"])
    assert doc["type"] == "WithTarget" and doc["result"] == {"astar_result_type": "FoundTarget", "cost": 2}
    assert doc["sequences"]["reference_name"] == "ref " and doc["template_switch_amount"] == 1.0
    # the written alignment rescoring to the written cost under the reference cost function
    flat = oracle.FlatConfig(tsa_config.parse(workloads.sample_config_text(), "dna-n"))
    ops = ops_from_toml(doc["alignment"])
    cost, er, eq, ok = oracle.rescore(flat, doc["sequences"]["reference"], doc["sequences"]["query"], ops)
    assert ok and cost == 2 and (er, eq) == (44, 44)
    entrance = doc["alignment"][1]
    assert entrance[0] == 5 and entrance[1]["TemplateSwitchEntrance"]["first_offset"] == -4  # |offset| + 1 (reverse)


def test_without_target_and_flags(workdir, toml_golden):
    r = run(workdir, "align", "-p", "test_files/twin_100_0.01.fa", "--cost-limit", "0", "--no-ts", "-o", "none.toml")
    assert r.returncode == 0
    assert r.stdout.splitlines()[:2] == ["No alignment found", "Exceeded cost limit of 0"]
    doc = tomllib.loads((workdir / "none.toml").read_text())
    golden = toml_golden["twin_ari_chrX_146823507_146823598_no_ts.toml"]["parsed"]
    assert set(doc) == set(golden) and doc["result"] == {"astar_result_type": "ExceededCostLimit", "cost_limit": 0}
    # heuristic flags of the reference are accepted and ignored; ranges; skip characters; separate files
    r = run(workdir, "align", "-p", "test_files/twin_10_ts.fa", "--ts-min-length-strategy", "none", "--ts-total-length-strategy=none",
            "--ts-node-ord-strategy", "anti-diagonal", "--rq-ranges", "R2..8Q2..8", "--skip-characters", "-", "-l", "debug", "--dont-extend-beyond-range")
    assert r.returncode == 0 and "Reference offset: 2" in r.stdout
    # by default the alignment is extended beyond the range while the cost does not increase (alignment_result.rs:247-395):
    # here over the two matching characters on either side, so the reported offsets move to 0
    r = run(workdir, "align", "-p", "test_files/twin_10_ts.fa", "--rq-ranges", "R2..8Q2..8", "--skip-characters", "-")
    assert r.returncode == 0 and "Reference offset: 0" in r.stdout and r.stdout.splitlines()[0].startswith("CIGAR: 2=[TS") and r.stdout.splitlines()[0].endswith("]2=")
    r = run(workdir, "align", "-r", "test_files/reference_a.fa", "-q", "test_files/query_a.fa", "--no-ts")
    assert r.returncode == 0 and "CIGAR: " in r.stdout


def test_errors(workdir):
    assert run(workdir, "align").returncode != 0                                   # no input
    assert run(workdir, "align", "-p", "test_files/missing.fa").returncode != 0
    assert run(workdir, "align", "-p", "test_files/twin_a.fa", "-a", "klingon").returncode == 2
    assert run(workdir, "align", "-p", "test_files/twin_a.fa", "--bogus").returncode == 2
    assert run(workdir, "show", "-i", "x.toml").returncode == 2
    (workdir / "bad.fa").write_text(">a\nACGX\n>b\nACGT\n")
    r = run(workdir, "align", "-p", "bad.fa")
    assert r.returncode == 1 and "non-alphabet character" in r.stderr
    r = run(workdir, "align", "-p", "test_files/twin_embedded.fa", "--use-embedded-rq-ranges", "--rq-ranges", "R0..1Q0..1")
    assert r.returncode == 1 and "Redundant" in r.stderr
    r = run(workdir, "align", "-p", "test_files/twin_a.fa", "-c", "nowhere")
    assert r.returncode == 1 and "config" in r.stderr


def test_batch_front_end(workdir, pairs):
    # --pairs: a multi-FASTA of several pairs (records 2k, 2k+1) and a TSV go through ONE tsa_align_batch call; every pair gets the
    # result the single-pair CLI gives it (the reference's own front-end accepts exactly two records, fasta_parser.rs:157-173)
    import json
    names = ("twin_a.fa", "twin_show_ts_indel1.fa", "twin_10_ts.fa")
    (workdir / "batch.fa").write_text("".join(pairs[n]["raw"] if pairs[n]["raw"].endswith("\n") else pairs[n]["raw"] + "\n" for n in names))
    os.makedirs(workdir / "out_dir", exist_ok=True)
    r = run(workdir, "align", "--pairs", "batch.fa", "-o", "out_dir", "--output-jsonl", "out.jsonl")
    assert r.returncode == 0, r.stderr
    rows = [ln.split("\t") for ln in r.stdout.splitlines()]
    assert len(rows) == len(names) and [row[0] for row in rows] == ["0", "1", "2"]
    recs = [json.loads(ln) for ln in (workdir / "out.jsonl").read_text().splitlines()]
    for k, name in enumerate(names):
        single = run(workdir, "align", "-p", f"test_files/{name}", "-o", f"single_{k}.toml")
        assert single.returncode == 0, single.stderr
        cigar = single.stdout.splitlines()[0].split(": ", 1)[1]
        assert rows[k][3] == "FoundTarget" and rows[k][6] == cigar and recs[k]["cigar"] == cigar
        a, b = tomllib.loads((workdir / "out_dir" / f"{k}.toml").read_text()), tomllib.loads((workdir / f"single_{k}.toml").read_text())
        for key in ("duration_seconds",):
            a.pop(key); b.pop(key)
        assert a == b and recs[k]["cost"] == a["result"]["cost"]
    # TSV input, --no-ts, a pair with a character outside the alphabet is reported and the rest of the batch still answered
    (workdir / "batch.tsv").write_text("p0\tACGTACGTAC\tACGTTCGTAC\n# comment\nACGT\tACG\np2\tACGZ\tACGT\n")
    r = run(workdir, "align", "--pairs", "batch.tsv", "--no-ts")
    rows = [ln.split("\t") for ln in r.stdout.splitlines()]
    assert r.returncode == 1 and len(rows) == 3
    assert rows[0][1] == "p0" and rows[0][3:6] == ["FoundTarget", "2", "0"] and rows[0][6] == "4=1X5="
    assert rows[1][1] == "pair1" and rows[1][3] == "FoundTarget" and rows[2][3] == "Error"
    odd = run(workdir, "align", "--pairs", "test_files/reference_a.fa")
    assert odd.returncode != 0 and "even number of records" in odd.stderr


def test_toml_bytes_against_golden_no_ts(workdir, toml_golden, configs):
    # Byte-level layout of the TOML result file against the reference's committed --no-ts result files (the TS files need the GPU:
    # tests/test_gpu_parity.py::test_cli_toml_bytes_against_golden_gpu): every line that does not depend on the search statistics or
    # on the tie-break among optimal alignments is identical text, in the same order.
    from helpers import parse_config_any
    volatile = ("duration_seconds", "opened_nodes", "closed_nodes", "suboptimal_opened_nodes", "suboptimal_opened_nodes_ratio", "runtime", "memory", "alignment")
    checked = 0
    for name, g in toml_golden.items():
        p = g["parsed"]
        if p["type"] != "WithTarget" or "no_ts" not in name:
            continue
        seqs = p["sequences"]
        ocfg = parse_config_any(configs[g["config"]])
        flat = oracle.FlatConfig(ocfg)
        _, er, eq, ok = oracle.rescore(flat, seqs["reference"], seqs["query"], ops_from_toml(p["alignment"]), p["reference_offset"], p["query_offset"])
        assert ok
        d = workdir / ("bytes_" + name.replace(".toml", ""))
        os.makedirs(d / "cfg", exist_ok=True)
        (d / "cfg" / "config.tsa").write_text(configs[g["config"]])
        (d / "pair.fa").write_text(f">{seqs['reference_name'].rstrip(' ')}\n{seqs['reference']}\n>{seqs['query_name'].rstrip(' ')}\n{seqs['query']}\n")
        r = run(d, "align", "-p", "pair.fa", "-c", "cfg", "-a", ocfg.alphabet, "-o", "out.toml", "--dont-extend-beyond-range", "--no-ts",
                "--rq-ranges", f"R{p['reference_offset']}..{er}Q{p['query_offset']}..{eq}")
        assert r.returncode == 0, (name, r.stderr)
        ours, gold = (d / "out.toml").read_text().splitlines(), g["raw"].splitlines()
        assert [ln.split(" =")[0] for ln in ours] == [ln.split(" =")[0] for ln in gold], name
        for a, b in zip(ours, gold):
            if a.split(" =")[0] not in volatile:
                assert a == b, (name, a[:120], b[:120])
        checked += 1
    assert checked >= 3
