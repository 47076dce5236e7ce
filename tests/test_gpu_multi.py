"""Multi-GPU: clusters are independent, so the host binary shards batches of clusters dynamically over all
devices (one worker thread + context per device, host-side gather of per-cluster files, no collective)."""
import json
import os
import subprocess

import pytest

from hylight_b200 import api, synth
from oracle import orc

from util import small_cluster, to_orc_batch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_binary_shards_clusters_over_all_devices(tmp_path):
    n_dev = api.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    exe = os.path.join(ROOT, "hylight_b200", "hylight-ovlp")
    lines, batches = [], []
    for c in range(4 * n_dev):
        b = small_cluster(300 + c, n_fragments=400 + 50 * (c % 3))
        synth.write_fasta(b, 0, str(tmp_path / ("c%d.fa" % c)))
        lines.append("%s\t%s" % (tmp_path / ("c%d.fa" % c), tmp_path / ("c%d.txt" % c)))
        batches.append(b)
    (tmp_path / "list.tsv").write_text("\n".join(lines) + "\n")
    r = subprocess.run([exe, "--clusters", str(tmp_path / "list.tsv"), "--devices", "all", "--batch_bases", "1",
                        "--workers_per_device", "1", "--stats", str(tmp_path / "stats.jsonl")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for c, b in enumerate(batches):
        rec, blk, coff, st = orc.overlap_batch(to_orc_batch(b))
        ref = tmp_path / ("c%d.ref" % c)
        orc.write_overlaps(rec, str(ref))
        assert open(tmp_path / ("c%d.txt" % c)).read() == open(ref).read()
    stats = [json.loads(l) for l in open(tmp_path / "stats.jsonl")]
    summary = [s for s in stats if s.get("summary")]
    stats = [s for s in stats if not s.get("summary")]
    assert len(summary) == 1 and summary[0]["clusters"] == len(batches) and summary[0]["records"] == sum(s["records"] for s in stats)
    assert len(stats) == len(batches)
    assert len({s["device"] for s in stats}) >= 2  # the dynamic queue fed more than one GPU


def test_every_device_gives_identical_records():
    n_dev = api.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    b = small_cluster(350, n_fragments=600)
    outs = []
    for d in range(n_dev):
        with api.Context(d) as c:
            outs.append(c.overlap_host(b)[0])
    for o in outs[1:]:
        assert (o == outs[0]).all()
