"""Regenerates the committed fixtures under tests/golden/ from the reference checkout.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Sources (reference file:line):
  fermion_ref.npy      <- tests/data_ref/fermion_ref.txt (256 complex amplitudes, tests/operator_test.py:242)
  reference_goldens.json:
     rbm_weights       <- tests/sampler_test.py:35-39 (== tests/tdvp_test.py:65-69)
     zz_trajectory     <- tests/tdvp_test.py:122-127
     gs_energies       <- tests/tdvp_test.py:30-32, tests/minsr_test.py:18-20
     fermion_energy    <- tests/operator_test.py:254
"""
import json
import os
import re
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def grab_floats(path, first, last):
    with open(path) as f:
        lines = f.readlines()[first - 1:last]
    return [float(x) for x in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", "".join(lines))]


def main():
    b = np.loadtxt(os.path.join(REF, "tests/data_ref/fermion_ref.txt"), dtype=np.complex128)
    np.save(os.path.join(HERE, "fermion_ref.npy"), b)
    w = grab_floats(os.path.join(REF, "tests/sampler_test.py"), 36, 38)
    zz = grab_floats(os.path.join(REF, "tests/tdvp_test.py"), 123, 126)
    assert len(w) == 16 and len(zz) == 21, (len(w), len(zz))
    out = {
        "rbm_weights": w,
        "zz_trajectory": zz,
        "gs_hx": [-1.3, -0.3],
        "gs_energies": [-6.10160339, -4.09296160],
        "fermion_energy": -9.95314531,
    }
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
