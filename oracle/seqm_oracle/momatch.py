"""MO crossing matcher -- restatement of Energy._crossing_match_molecular_orbitals / _grouped
(seqm/basics.py:596-719), which the reference runs on every forward after the first one on the same Molecule
(basics.py:846-857).  TEST INFRASTRUCTURE ONLY (see seqm_oracle/__init__.py).

Per molecule, separately for the occupied block [0, nocc) and the virtual block [nocc, norb):
  S = |C_old^T C_new|                                                  (basics.py:639-641)
  p[k] = argmax_l S[k, l]; if p is not a permutation, repair greedily:  (basics.py:643-655)
      rows are served in the order of descending margin (best minus second best entry of the row; the single
      entry for a 1x1 block), each taking its largest still unused column  (greedy_unique_perm, basics.py:606-626)
  C_out[:, k] = sign(<C_new[:, p[k]], C_old[:, k]>) C_new[:, p[k]], sign(0) = +1     (basics.py:628-635, 661-662)
  e_out[k] = e[p[k]]                                                   (basics.py:666-669)
Orbitals beyond norb (padding of a mixed batch) are left untouched (basics.py:699-717)."""
import numpy as np


def _block_perm(S):
    r = S.shape[0]
    p = np.argmax(S, axis=1)
    if sorted(p.tolist()) == list(range(r)):
        return p
    pref = np.argsort(-S, axis=1, kind="stable")
    if r > 1:
        top2 = np.take_along_axis(S, pref[:, :2], axis=1)
        prio = top2[:, 0] - top2[:, 1]
    else:
        prio = S[:, 0]
    p = np.empty(r, dtype=np.int64)
    used = np.zeros(r, dtype=bool)
    for row in np.argsort(-prio, kind="stable").tolist():
        col = next(c for c in pref[row].tolist() if not used[c])
        p[row] = col
        used[col] = True
    return p


def match_orbitals(V_new, V_old, nocc, norb, e):
    """V_new, V_old: (nmol, nmax, nmax), column = MO; nocc, norb: (nmol,); e: (nmol, >= nmax).
    Returns (V_out, e_out) like the reference's grouped matcher."""
    V_new, V_old, e = np.asarray(V_new), np.asarray(V_old), np.asarray(e)
    V_out, e_out = V_new.copy(), e.copy()
    for m in range(V_new.shape[0]):
        n, no = int(norb[m]), int(nocc[m])
        for lo, hi in ((0, no), (no, n)):
            if hi <= lo:
                continue
            Cn, Co = V_new[m, :n, lo:hi], V_old[m, :n, lo:hi]
            p = _block_perm(np.abs(Co.T @ Cn))
            Cp = Cn[:, p]
            s = np.sign((Cp * Co).sum(axis=0))
            s[s == 0] = 1.0
            V_out[m, :n, lo:hi] = Cp * s
            e_out[m, lo:hi] = e[m, lo:hi][p]
    return V_out, e_out
