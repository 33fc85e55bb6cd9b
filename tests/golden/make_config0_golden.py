"""BASELINE configs[0] golden: the reference binary's `simuvars` output for (synthetic chr20, testData snp.txt, testData vars.txt).
Only its SHA-256 and size are committed (the FASTA is 128 MB). Run in the build container: python tests/golden/make_config0_golden.py"""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    ref, snp, var = H.config0_inputs(d)
    r = H.run_reference_simuvars(ref, snp, var, os.path.join(d, "cell.fa"))
    assert r.returncode == 0, r.stderr.decode()
    data = H.read_bytes(os.path.join(d, "cell.fa"))
    recs = H.read_fasta_records(os.path.join(d, "cell.fa"))
    json.dump({"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data), "records": [(n, int(len(s))) for n, s in recs],
               "generator": "oracle/_ref/bin/scssim simuvars (reference binary) on helpers.config0_inputs()"},
              open(os.path.join(HERE, "testdata", "config0.json"), "w"), indent=1)
    print(open(os.path.join(HERE, "testdata", "config0.json")).read())
