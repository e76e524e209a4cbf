"""Generate the committed fixtures under tests/golden/ from the read-only reference.

Run in the BUILD container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes
  hbs.npz / goku.npz  raw in-repo arrays (reference ``data/50_LR_3_HR`` and
                      ``data/matter_power_1128_..._z0``; read exactly as
                      ``mfgpflow/data_loader.py:288-322`` does with np.loadtxt) plus the
                      KMeans(random_state=42) inducing-point centres the reference
                      constructors compute (``singlebin_svgp.py:50-51``).
  goldens.json        the known answers G1-G7 scraped from the saved cell outputs of the
                      reference notebooks, with their provenance.
"""
import json
import os
import re

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

DATASETS = {
    "hbs": ("data/50_LR_3_HR", 50),
    "goku": ("data/matter_power_1128_Box1000_Part750_36_Box1000_Part3000_z0", 300),
}


def load_raw(folder):
    f = lambda n: np.loadtxt(os.path.join(REF, folder, n))
    return dict(
        X_LF=f("train_input_fidelity_0.txt"),
        X_HF=f("train_input_fidelity_1.txt"),
        Y_LF=f("train_output_fidelity_0.txt"),
        Y_HF=f("train_output_fidelity_1.txt"),
        X_test=np.atleast_2d(f("test_input.txt")),
        Y_test=np.atleast_2d(f("test_output.txt")),
        input_limits=f("input_limits.txt"),
        kf=f("kf.txt"),
    )


def kmeans_centres(raw, M):
    from sklearn.cluster import KMeans

    lim = raw["input_limits"]
    unit = lambda x: (x - lim[:, 0]) / (lim[:, 1] - lim[:, 0])
    XL, XH = unit(raw["X_LF"]), unit(raw["X_HF"])
    X = np.vstack([np.hstack([XL, np.zeros((len(XL), 1))]), np.hstack([XH, np.ones((len(XH), 1))])])
    return KMeans(n_clusters=M, random_state=42).fit(X).cluster_centers_


def scrape(nb, cell, pattern, count=1):
    d = json.load(open(os.path.join(REF, "notebooks", nb)))
    text = ""
    for o in d["cells"][cell].get("outputs", []):
        if "text" in o:
            text += "".join(o["text"])
    vals = [float(m) for m in re.findall(pattern, text)]
    return vals[:count]


def main():
    for name, (folder, M) in DATASETS.items():
        raw = load_raw(folder)
        raw[f"Z_kmeans{M}"] = kmeans_centres(raw, M)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **raw)
        print(name, {k: v.shape for k, v in raw.items()})

    it = r"Iteration \d+: (?:Loss|ELBO) = (-?[0-9.eE+-]+)"
    g = {}
    mp = scrape("demo: matter power.ipynb", 3, it, 10)
    g["G1_hbs_gpr_lml_init"] = dict(value=mp[0], source="notebooks/demo: matter power.ipynb cell 3, 'Iteration 0'")
    g["G3_hbs_gpr_lml_adam_traj"] = dict(
        iters=[100, 200, 300, 400, 500], values=mp[1:6], later_iters=[600, 700, 800, 900], later_values=mp[6:10],
        source="same cell, Adam(lr=0.1); printed value is LML evaluated BEFORE the update of that iteration",
    )
    gk = scrape("demo: goku power spectra.ipynb", 4, it, 10)
    g["G2_goku_gpr_lml_init"] = dict(value=gk[0], source="notebooks/demo: goku power spectra.ipynb cell 4")
    g["G2b_goku_gpr_lml_adam_traj"] = dict(iters=[100, 200, 300, 400, 500], values=gk[1:6], source="same cell")
    g["G4_hbs_singlebin_negelbo_step1"] = dict(
        value=scrape("demo matter power single bin.ipynb", 3, it)[0],
        source="notebooks/demo matter power single bin.ipynb cell 3: -ELBO printed AFTER one Adam(0.1) step",
    )
    g["G5_goku_singlebin_negelbo_step1"] = dict(
        value=scrape("demo: goku power spectra.ipynb", 10, it)[0], source="goku nb cell 10"
    )
    g["G6_hbs_latent_negelbo_step1"] = dict(
        value=scrape("demo: matter power latent inference.ipynb", 4, it)[0],
        source="notebooks/demo: matter power latent inference.ipynb cell 4 (historical ctor: q_sqrt=0.1*I, L=10, M=50)",
    )
    g["G7_goku_latent_negelbo_step1"] = dict(
        value=scrape("demo: goku power spectra.ipynb", 22, it)[0], source="goku nb cell 22 (L=15, M=300)"
    )
    json.dump(g, open(os.path.join(HERE, "goldens.json"), "w"), indent=1)
    print(json.dumps(g, indent=1))


if __name__ == "__main__":
    main()
