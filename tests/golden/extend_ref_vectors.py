"""Copies further tables of the reference's hoomd/md/pytest/forces_and_energies.json into
tests/golden/ref_vectors.json (run in the build container, where /root/reference exists):

    python tests/golden/extend_ref_vectors.py Buckingham ExpandedLJ ExpandedGaussian LJGauss TWF
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/hoomd/md/pytest/forces_and_energies.json"

if __name__ == "__main__":
    with open(REF) as f:
        fe = json.load(f)
    path = os.path.join(HERE, "ref_vectors.json")
    with open(path) as f:
        out = json.load(f)
    for name in sys.argv[1:]:
        out["forces_and_energies"][name] = fe[name]
    with open(path, "w") as f:
        json.dump(out, f)
