"""The oracle reproduces the committed golden vectors (tests/golden/golden_small.npz, made by make_golden.py)."""
import os

import numpy as np

from oracle import gp_oracle as go

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.npz")


def cases():
    z = np.load(GOLD)
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


def test_oracle_reproduces_golden():
    z, names = cases()
    assert len(names) == 6
    for nm in names:
        X = np.ascontiguousarray(z[f"{nm}/X"].T)
        th, kind = z[f"{nm}/theta"], str(z[f"{nm}/kind"])
        for k in range(z[f"{nm}/Y"].shape[0]):
            r = go.eval_mll(X, z[f"{nm}/Y"][k], th, kind=kind, return_state=True)
            assert r["info"] == z[f"{nm}/info"][k]
            np.testing.assert_allclose(r["mll"], z[f"{nm}/mll"][k], rtol=1e-11)
            np.testing.assert_allclose(r["grad"], z[f"{nm}/grad"][k], rtol=1e-8, atol=1e-9)
            mu, var = go.predict(X, th, r["state"], np.ascontiguousarray(z[f"{nm}/Xtest"].T), kind=kind)
            np.testing.assert_allclose(mu, z[f"{nm}/mu"][k], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(var, z[f"{nm}/var"][k], rtol=1e-9, atol=1e-14)
