"""GPU parity of `resample` and the jump entry points around it (src/opensys/trajectory_jump.jl:53-56, :64-88): several
Liouvillians in `opensys` / `opensys_eig`, each proposing one operator with its own uniform, the winner drawn with the
first-power weights ‖Lψ‖.  Indices must match the oracle EXACTLY; applied states to 1e-12."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import fixtures as F
from oracle import oqb_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    import oqb_b200
    oqb_b200.get_context()
    return oqb_b200


def gamma(x):
    return x + 1 if x >= 0 else (1 - x) * np.exp(x)


def rand_c(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def test_lindblad_jump_two_liouvillians_resample(Q):
    from oqb_b200 import host as Hm
    rng = np.random.default_rng(11)
    q, n, B = 4, 16, 40
    Hd = sps.csc_matrix(-F.standard_driver(q))
    Hp = sps.csc_matrix(F.random_ising(q, rng))
    Z = [sps.csc_matrix(F.q_translate("".join("Z" if j == k else "I" for j in range(q)))) for k in range(q)]
    Xs = [sps.csc_matrix(F.q_translate("".join("X" if j == k else "I" for j in range(q)))) for k in range(2)]
    fs = [lambda s: 1 - s, lambda s: s]
    gz = [lambda s, k=k: 0.05 * (1 + k) for k in range(q)]
    gx = [lambda s, k=k: 0.3 + 0.1 * k * s for k in range(2)]
    Hg, Ho = Q.SparseHamiltonian(fs, [Hd, Hp], unit="ħ"), O.SparseHamiltonian(fs, [Hd, Hp], unit="ħ")
    lg = [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gz, Z)]), Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gx, Xs)])]
    lo = [O.LindbladLiouvillian([O.Lindblad(g, L.toarray()) for g, L in zip(gz, Z)]),
          O.LindbladLiouvillian([O.Lindblad(g, L.toarray()) for g, L in zip(gx, Xs)])]
    opg = Q.DiffEqLiouvillian(Hg, [], lg, n)
    opo = O.DiffEqLiouvillian(Ho, [], lo, n)
    psi = rand_c(rng, B, n)
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    uni = rng.random((3, B))
    mask = (rng.random(B) < 0.8).astype(np.uint8)
    pg, po = Q.ODEParams(None, 2.0), O.ODEParams(None, 2.0)
    dpsi = Q.DeviceBatch.from_host(psi[:, :, None].copy())
    pick, choices = Hm.lindblad_jump(opg, dpsi, pg, 1.0, uni, mask, apply=True)
    out = dpsi.to_host().reshape(B, -1)
    for b in range(B):
        if not mask[b]:
            assert pick[b] == -1 and np.array_equal(out[b], psi[b])
            continue
        win, ch, L = O.lindblad_jump_multi(opo, psi[b], po, 1.0, [uni[0, b], uni[1, b], uni[2, b]])
        assert pick[b] == win and list(choices[:, b]) == ch, b
        ref = L @ psi[b]
        ref /= np.linalg.norm(ref)
        assert np.linalg.norm(out[b] - ref) < 1e-12


def test_ame_jump_two_davies_generators_resample(Q):
    from oqb_b200 import host as Hm
    rng = np.random.default_rng(12)
    n, B = 12, 24
    w = np.sort(np.r_[np.repeat([-1.0, 0.5], 2), np.arange(4) * 0.75 + 1.0, rng.standard_normal(4)])
    v = np.linalg.qr(rand_c(rng, n, n))[0]
    Hmat = v @ np.diag(w) @ v.conj().T
    cs = [(lambda m: m + m.conj().T)(rand_c(rng, n, n)) for _ in range(3)]
    g2 = lambda x: 0.4 * gamma(x) + 0.1
    Hg, Ho = Q.DenseHamiltonian([lambda s: 1.0], [Hmat], unit="ħ"), O.DenseHamiltonian([lambda s: 1.0], [Hmat], unit="ħ")
    zero = lambda x: 0.0
    opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(cs[:2], gamma, zero), Q.DaviesGenerator(cs[2:], g2, zero)], [], n)
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator(cs[:2], gamma, zero), O.DaviesGenerator(cs[2:], g2, zero)], [], n)
    psi = rand_c(rng, B, n)
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    uni = rng.random((3, B))
    dpsi = Q.DeviceBatch.from_host(psi[:, :, None].copy())
    pick, choices = Hm.lindblad_jump(opg, dpsi, Q.ODEParams(None, 1.0), 0.5, uni, None, apply=True, w=w, v=v)
    out = dpsi.to_host().reshape(B, -1)
    for b in range(B):
        win, ch, L = O.lindblad_jump_multi(opo, psi[b], O.ODEParams(None, 1.0), 0.5, [uni[0, b], uni[1, b], uni[2, b]], w, v)
        assert pick[b] == win and list(choices[:, b]) == ch, b
        ref = L @ psi[b]
        ref /= np.linalg.norm(ref)
        assert np.linalg.norm(out[b] - ref) < 1e-12


def test_ame_jump_const_davies(Q):
    """ame_jump(::ConstDaviesGenerator) (:53-56) on the operators build_const_davies precomputes (src/opensys/davies.jl:173-209)."""
    from oqb_b200 import extra_liouvillians as X
    rng = np.random.default_rng(13)
    l, B = 8, 30
    w = np.array([-2.0, -2.0, -0.5, 0.3, 0.3, 1.1, 1.9, 2.7])
    cs = [(lambda m: m + m.conj().T)(rand_c(rng, l, l)) for _ in range(2)]
    lamb = lambda x: 0.1 * x
    Dg = X.build_const_davies(cs, w, gamma, lamb)
    Do = O.build_const_davies(cs, O.build_gap_indices(w, 8, 8), gamma, lamb)
    assert len(Dg.Linds) == len(Do.Linds)
    psi = rand_c(rng, B, l)
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    uni = rng.random(B)
    dpsi = Q.DeviceBatch.from_host(psi[:, :, None].copy())
    choice, prob = Dg.ame_jump(dpsi, uni, None, apply=True)
    choice, prob = choice.cpu().numpy(), prob.cpu().numpy()
    out = dpsi.to_host().reshape(B, -1)
    for b in range(B):
        k, pr = O.ame_jump_const(Do, psi[b], uni[b])
        assert choice[b] == k
        assert np.allclose(prob[b], pr, rtol=1e-12, atol=1e-14)
        ref = Do.Linds[k] @ psi[b]
        ref /= np.linalg.norm(ref)
        assert np.linalg.norm(out[b] - ref) < 1e-12
