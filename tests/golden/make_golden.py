#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the reference's cached solutions.

Run in the BUILD container only (needs /root/reference; the GPU box has neither it
nor any need for it -- the .npz files are committed).

Source: /root/reference/docs/data/<name>_<githash>.jld2, written by ``cached_solve!``
(/root/reference/src/docs_cache.jl:45-56,180-218): a serialized converged
``NamedTrajectory``.  JLD2 is an HDF5 dialect and no HDF5 reader is installed, so the
``datavec`` Float64 block (D*K contiguous little-endian doubles, knot-major) is located by
searching the file for the byte image of the known first knot's state (the ``initial``
constraint of the problem) and validated structurally (finite, dt row constant / positive,
t row = cumsum(dt)).

Each fixture stores Z (D x K, Fortran order), the system definition needed to rebuild the
generator with oracle/systems.py, and the docs script that produced it.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
DATA = "/root/reference/docs/data"

from oracle import isomorphisms as iso  # noqa: E402

SPECS = {
    # name: (file, kind, d, m, K, first-knot state, docs source)
    "two_qubit_zoh": dict(
        file="two_qubit_zoh_57a874f.jld2", kind="unitary", d=4, m=4, K=200,
        x0=iso.operator_to_iso_vec(np.eye(4)),
        system="MultiTransmonSystem([4.0,4.1],[0.2,0.2],[[0,.1],[.1,0]]; levels_per_transmon=2, drive_bounds=0.1)",
        source="docs/literate/two_qubit_gate_validation.jl:51-56,136-149", tight=True),
    "trajectories_density": dict(
        file="trajectories_density_88fab3e.jld2", kind="density", d=2, m=2, K=50,
        x0=iso.density_to_compact_iso(np.diag([1.0, 0.0])),
        system="OpenQuantumSystem(Z,[X,Y],[1,1]; dissipation_operators=[[0.1 0;0 0]])",
        source="docs/literate/concepts/trajectories.jl:130-153", tight=True),
    "systems_cat_density": dict(
        file="systems_cat_density_ce912f0.jld2", kind="density", d=6, m=2, K=11,
        x0=iso.density_to_compact_iso(np.diag([1.0, 0, 0, 0, 0, 0])),
        system="CatSystem(cat_levels=3, buffer_levels=2)",
        source="docs/literate/systems/cat_qubits.jl:118-148", tight=True),
    "trajectories_ket": dict(
        file="trajectories_ket_573ffb2.jld2", kind="ket", d=2, m=2, K=100,
        x0=iso.ket_to_iso(np.array([1.0, 0.0])),
        system="QuantumSystem(Z,[X,Y],[1,1])",
        source="docs/literate/concepts/trajectories.jl:38-83", tight=True),
    "trajectories_unitary": dict(
        file="trajectories_unitary_573ffb2.jld2", kind="unitary", d=2, m=2, K=100,
        x0=iso.operator_to_iso_vec(np.eye(2)),
        system="QuantumSystem(Z,[X,Y],[1,1])",
        source="docs/literate/concepts/trajectories.jl:38-57", tight=False),
    # MultiKetTrajectory: two kets sharing one generator and one control row; knot column
    # [psi1 (4) | psi2 (4) | dt | t | u | du | ddu]  (one BilinearIntegrator per state,
    # src/control/integrators.jl:102-117).  Stopped at max_iter = 50: a valid input, not a delta ~ 0 KAT.
    "trajectories_multi": dict(
        file="trajectories_multi_573ffb2.jld2", kind="multiket", d=2, m=2, K=100, n_states=2,
        x0=np.concatenate([iso.ket_to_iso(np.array([1.0, 0.0])), iso.ket_to_iso(np.array([0.0, 1.0]))]),
        system="QuantumSystem(Z,[X,Y],[1,1]); MultiKetTrajectory(sys, pulse, [|0>,|1>], [|1>,|0>])",
        source="docs/literate/concepts/trajectories.jl:86-113", tight=False),
}


def n_x_of(kind, d):
    return {"ket": 2 * d, "unitary": 2 * d * d, "density": d * d}[kind]


def extract(spec):
    raw = open(os.path.join(DATA, spec["file"]), "rb").read()
    n_x = 2 * spec["d"] * spec["n_states"] if spec["kind"] == "multiket" else n_x_of(spec["kind"], spec["d"])
    D = n_x + 2 + 3 * spec["m"]
    K = spec["K"]
    pat = np.asarray(spec["x0"], dtype="<f8").tobytes()
    start = 0
    while True:
        i = raw.find(pat, start)
        if i < 0:
            raise RuntimeError("datavec not found in " + spec["file"])
        start = i + 8
        if i + 8 * D * K > len(raw):
            continue
        Z = np.frombuffer(raw, dtype="<f8", count=D * K, offset=i).reshape(D, K, order="F")
        dt, t = Z[n_x], Z[n_x + 1]
        if not np.all(np.isfinite(Z)) or np.any(dt <= 0) or np.any(dt > 10):
            continue
        if abs(t[0]) > 1e-12 or np.max(np.abs(np.cumsum(dt[:-1]) - t[1:])) > 1e-6:
            continue
        return np.array(Z, order="F"), i


def main():
    meta = {}
    for name, spec in SPECS.items():
        Z, off = extract(spec)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), Z=Z)
        meta[name] = {k: v for k, v in spec.items() if k != "x0"}
        meta[name].update(D=int(Z.shape[0]), byte_offset=int(off))
        print(name, Z.shape, "offset", off)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
