"""Extract the GaussianMixture parameters of the reference's UPS instance distribution
(gaussian_mixture/cvrptw.joblib, cvrp.joblib: sklearn 0.21.3 pickles) WITHOUT importing the
pickled sklearn classes: a restricted unpickler maps every sklearn class to a plain attribute bag and
allows numpy reconstruction only.  Run once in the build container (needs /root/reference):

    python tools/extract_gmm.py

Writes mit_rl-vrptw_b200/data/gmm_{cvrptw,cvrp}.npz (weights [K], means [K,d], covariances [K,d,d]).
These are data fixtures (distribution parameters), not code.
"""
import io
import os
import pickle
import sys

import joblib.numpy_pickle as jnp
import numpy as np

REF = "/root/reference/gaussian_mixture"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mit_rl-vrptw_b200", "data")


class Bag:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, st):
        self.__dict__.update(st)


class Restricted(jnp.NumpyUnpickler):
    def find_class(self, module, name):
        if module.startswith("sklearn"):
            return Bag
        if module.startswith(("numpy", "joblib")) or (module, name) in {("builtins", "dict"), ("collections", "OrderedDict")}:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"blocked {module}.{name}")


def load(path):
    with open(path, "rb") as f:
        buf = io.BytesIO(f.read())
    up = Restricted(path, buf, ensure_native_byte_order=True) if "ensure_native_byte_order" in jnp.NumpyUnpickler.__init__.__code__.co_varnames \
        else Restricted(path, buf)
    return up.load()


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ("cvrptw", "cvrp"):
        g = load(os.path.join(REF, name + ".joblib"))
        w, mu, cov = np.asarray(g.weights_), np.asarray(g.means_), np.asarray(g.covariances_)
        print(name, "K", w.shape, "means", mu.shape, "cov", cov.shape, "type", getattr(g, "covariance_type", "?"))
        np.savez(os.path.join(OUT, f"gmm_{name}.npz"), weights=w, means=mu, covariances=cov)


if __name__ == "__main__":
    sys.exit(main())
