"""Builds tests/golden/lung.npz from the reference's bundled example data (run in the build container, where
/root/reference exists; the GPU box only sees the committed .npz).  Data files are fixtures, not reference source code.

  python tests/golden/make_lung_fixture.py
"""
import os
import numpy as np

REF = "/root/reference/data"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    names, rows = [], []
    with open(os.path.join(REF, "example_lung.tpm")) as f:
        next(f)
        for line in f:
            p = line.split()
            names.append(p[0]); rows.append([float(x) for x in p[1:]])
    gl_names, gl = [], []
    with open(os.path.join(REF, "example_gene_length.txt")) as f:
        next(f)
        for line in f:
            p = line.split()
            gl_names.append(p[0]); gl.append(float(p[1]))
    assert names == gl_names
    tpm = np.array(rows)
    # docs/tutorial.html:681-691 — P(active) of the first 10 genes from the tutorial run (one unseeded chain)
    pinned = np.array([0.955, 0.094, 0.901, 1.000, 0.000, 0.974, 0.724, 1.000, 1.000, 1.000])
    np.savez_compressed(os.path.join(HERE, "lung.npz"), tpm=tpm, gene_length=np.array(gl), gene_names=np.array(names),
                        tutorial_prob_active_first10=pinned, threshold_i=0.31, threshold_a=np.array([0.31, 5.39]))
    print("wrote lung.npz", tpm.shape, "zeros per lib", (tpm == 0).sum(0), "all-zero genes", int((tpm == 0).all(1).sum()))


if __name__ == "__main__":
    main()
