"""hylight-cluster (host half) against the unmodified reference scripts
bin_pointer_limited_filechunks_shortpath2.py -> getclusters.py -> get_fq_cluster.py: same clusters, same FASTQ files.
Golden digests: tests/golden/clusters.json (tests/golden/make_golden_clusters.py ran the reference in the build container)."""
import hashlib
import importlib.util
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "hylight_b200", "hylight-cluster")
spec = importlib.util.spec_from_file_location("make_golden_clusters", os.path.join(ROOT, "tests", "golden", "make_golden_clusters.py"))
gold = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gold)
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "clusters.json")))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "hylight_b200", "csrc"), "../hylight-cluster"])
    d = str(tmp_path_factory.mktemp("clusters"))
    assert gold.inputs(d) == GOLDEN["paf_sha256"], "the input generator no longer reproduces the golden overlap file"
    return d


def run(d, case, tag, *extra):
    r = subprocess.run([EXE, "--paf", "shortr2.paf", "--readnames", "readnames.txt", "--size", case["size"], "--threads", str(case["threads"]),
                        "--json", tag, "--fastq", "reads.fq", "--outdir", "fq_" + tag] + list(extra), cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout)


@pytest.mark.parametrize("i", range(len(GOLDEN["cases"])))
def test_same_clusters_and_fastq_as_the_reference_scripts(workdir, i):
    case = GOLDEN["cases"][i]
    tag = "m%d" % i
    st = run(workdir, case, tag)
    grouped = json.load(open(os.path.join(workdir, tag + "_clusters_grouped.json")))
    assert gold.digest_grouped(grouped) == case["grouped"]
    assert json.load(open(os.path.join(workdir, tag + "_clustersizes.json"))) == case["clustersizes"]
    assert gold.digest_tree(os.path.join(workdir, "fq_" + tag)) == case["fastq"]
    assert st["clusters"] == len(case["grouped"]) and st["clustered_reads"] == sum(n for n, _ in case["grouped"].values())
    assert st["rounds"] == -(-st["chunks"] // case["threads"])


def test_keep_tail_restores_the_reads_getclusters_drops(workdir):
    """getclusters.py:55-61 slices the clustered reads into `threads` pieces of len // threads: the last len % threads reads
    appear in no cluster.  The tool does the same by default and says how many; --keep_tail keeps them."""
    case = GOLDEN["cases"][0]  # size 200, 2 threads: 2983 clustered reads, one dropped
    st = run(workdir, case, "tail0")
    st_all = run(workdir, case, "tail1", "--keep_tail")
    assert st["tail_reads_dropped"] == 1 and st_all["tail_reads_dropped"] == 0
    assert st_all["clustered_reads"] == st["clustered_reads"] + 1 and st_all["fastq_records"] == st["fastq_records"] + 2
    a = json.load(open(os.path.join(workdir, "tail0_clusters_grouped.json")))
    b = json.load(open(os.path.join(workdir, "tail1_clusters_grouped.json")))
    extra = {r for v in b.values() for r in v} - {r for v in a.values() for r in v}
    assert extra == {"frag2999"}


def test_chunking_is_part_of_the_result(workdir):
    """the reference filters a round's lines against the state at the START of the round, so the chunk size and the thread
    count change which merges happen: --chunk_bytes / --threads are semantics, not tuning"""
    case = dict(GOLDEN["cases"][0])
    one = run(workdir, dict(case, threads=1), "c1")
    big = run(workdir, dict(case, threads=1), "c2", "--chunk_bytes", str(1 << 30))
    assert big["chunks"] == 1 and one["chunks"] > 1
    assert one["lines_kept"] < big["lines_kept"]  # later rounds drop lines of clusters that are full or merged by then


def test_unknown_read_and_short_line_fail_loudly(tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("binary not built")
    (tmp_path / "names.txt").write_text("a\nb\n")
    (tmp_path / "x.paf").write_text("a/1\t1\t0\t1\t+\tzz/1\t1\t0\t1\t1\t1\t1\n")
    r = subprocess.run([EXE, "--paf", "x.paf", "--readnames", "names.txt", "--size", "10"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "zz" in r.stderr
    (tmp_path / "y.paf").write_text("a/1\t1\t0\n")
    r = subprocess.run([EXE, "--paf", "y.paf", "--readnames", "names.txt", "--size", "10"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "columns" in r.stderr
