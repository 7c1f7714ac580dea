#!/usr/bin/env python
"""Writes tests/golden/reference_known_answers.json.

The reference (OpenQuantumBase.jl) is pure Julia and Julia is not available in this image, so these fixtures are NOT
outputs of a reference run: they are the known-answer values the reference's OWN test-suite asserts for the hot path,
transcribed literal by literal with the test file:line each one comes from (paths relative to /root/reference).  Values
the reference states symbolically (e.g. exp(-0.5), 2πη/β) are written as the expression's float64 value.
Complex matrices are stored as [[re, im], …] row-major lists.

  python tests/golden/make_golden.py      # regenerates the JSON (deterministic, no inputs)
"""
import json
import math
import os

e = math.exp(-0.5)
a, b = math.sin(10) / 2, (1 - math.cos(10)) / 2


def cm(rows):
    return [[[complex(x).real, complex(x).imag] for x in r] for r in rows]


GOLD = {
    "lindblad_rhs_Lz_on_plus_x": {"source": "test/opensys/lindblad.jl:25-31", "gamma": [0.5], "ops": ["sz"],
                                   "state": "|+x><+x|", "t": 5.0, "tf": 5.0, "drho": cm([[0, -0.5], [-0.5, 0]])},
    "lindblad_cache_Lz": {"source": "test/opensys/lindblad.jl:33-35", "cache": cm([[-0.25, 0], [0, -0.25]])},
    "lindblad_rhs_Lz_Lx_on_plus_y": {"source": "test/opensys/lindblad.jl:38-42", "gamma": [0.5, 0.2], "ops": ["sz", "sx"],
                                      "state": "|+y><+y|", "drho": cm([[0, 0.7j], [-0.7j, 0]])},
    "lindblad_cache_Lz_Lx": {"source": "test/opensys/lindblad.jl:44-46", "cache": cm([[-0.35, 0], [0, -0.35]])},
    "dense_hamiltonian_half": {"source": "test/hamiltonian/dense_hamiltonian.jl:12,23-25", "s": 0.5,
                                "H": cm([[math.pi, math.pi], [math.pi, -math.pi]]),
                                "cache": cm([[-1j * math.pi, -1j * math.pi], [-1j * math.pi, 1j * math.pi]])},
    "redfield_lambda_code_form": {"source": "src/opensys/redfield.jl:46-64 with the model of test/opensys/redfield.jl:3-24 "
                                            "(U = exp(-i t sx), S = sz, C = 1, t = Ta = 5); closed form, SURVEY App. C",
                                  "Lambda": cm([[a, 1j * b], [-1j * b, -a]]), "drho_on_plus_x": cm([[0, -2 * a], [-2 * a, 0]]),
                                  "tolerance": 1e-6},
    "correlated_davies_zero": {"source": "test/opensys/davies.jl:173-184", "couplings": ["sigma_plus", "sigma_minus"],
                                "inds": [[1, 2], [2, 1]], "w": [1.0, 2.0], "rho": cm([[0.5, 0], [0, 0.5]]), "s": 0.5,
                                "du": cm([[0, 0], [0, 0]])},
    "correlated_davies_closed_form": {"source": "test/opensys/davies.jl:186-194", "couplings": ["sigma_minus", "sigma_minus"],
                                       "inds": [[1, 2], [2, 1]], "w": [1.0, 2.0], "rho": cm([[0.5, 0.5], [0.5, 0.5]]), "s": 0.5,
                                       "du": cm([[-e, -0.5 * e], [-0.5 * e, e]])},
    "gap_indices_example": {"source": "test/opensys/util.jl:3-9", "w": [-3, 1, 3, 3, 4.5, 5.5, 8], "digits": 8, "sigdigits": 8,
                             "note": "GapIndices(w) must list the same (gap, a, b) groups as find_unique_gap(w)"},
    "ohmic_gamma0": {"source": "test/coupling_bath_interaction/bath.jl:3-14", "eta": 1e-4, "fc": 4, "T": 16,
                      "rule": "gamma(0) == 2*pi*eta/beta"},
    "odeparams": {"source": "test/annealing.jl:19-22", "tf": 10.0, "t": 5.0, "s": 0.5},
    "ohmic_lambshift_at_zero": {"source": "test/coupling_bath_interaction/bath.jl:3-16", "eta": 1e-4, "wc": 8 * math.pi,
                                "beta": 1 / 2.23, "w": 0.0, "S": -0.0025132734115775254, "rtol": 1e-4, "atol": 1e-4},
    "lambshift_piecewise_exponential_spectrum": {"source": "test/coupling_bath_interaction/bath.jl:58-80",
                                                 "spectrum": "w >= 0 ? exp(-w) : exp(0.8w)", "w": 0.0, "S": 0.03551,
                                                 "atol": 1e-4, "table": {"lo": -5, "hi": 5, "length": 1000}},
    "bspline_linear_data": {"source": "test/interpolations.jl:3-22", "grid": {"lo": 0, "hi": 10, "length": 100},
                            "y": "10 - x", "test_grid": {"lo": 0, "hi": 10, "length": 50}, "outside": [[-1, 0.0], [11, 0.0]],
                            "gradient": -1.0},
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_known_answers.json")
    with open(out, "w") as f:
        json.dump(GOLD, f, indent=1, sort_keys=True)
    print(out, len(GOLD), "entries")
