"""Fixture of BASELINE config 1's molecules: the dimer2 (1A2K) receptor and ligand of examples/run_correlate.sh as
atom arrays (names, original coordinates, radii after prm assignment), read through the compiled reference's own
libmol2 calls.  Run in the build container:   python tests/golden/make_golden_dimer2.py
The ft files and the rotation set of that example are missing from the reference (.MISSING_LARGE_BLOBS); the seeded
generator libfmftsaxs_b200/workload.py:make_config1 stands in for them ("Config 1").
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refso  # noqa: E402

D = "/root/reference/examples/dimer2/"
PRM = "/root/reference/prms/atoms.0.0.6.prm.ms.3cap+0.5ace.Hr0rec"
out = {}
for key, fn in (("rec", "r_u_nmin.pdb"), ("lig", "l_u_nmin.pdb")):
    m = refso.load_pdb(D + fn, PRM, 0)
    out.update({key + "_xyz": m["xyz"], key + "_radius": m["radius"], key + "_res": np.array(m["res"], dtype="S8"),
                key + "_atm": np.array(m["atm"], dtype="S8")})
    print(key, len(m["res"]), "atoms")
np.savez_compressed(os.path.join(HERE, "golden_dimer2.npz"), **out)
