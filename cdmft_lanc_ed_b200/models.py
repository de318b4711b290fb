"""Synthetic impurity models = the *inputs* of the Hamiltonian-times-vector path.

What ED_HAMILTONIAN reads from module globals in the reference (``impHloc``
ED_VARS_GLOBAL.f90:119, ``Hbath_build(lambda)`` ED_BATH/hbath_setup.f90:240-250,
``dmft_bath%item(ib)%v``, interaction scalars ED_INPUT_VARS.f90:19-32) is packed
here in one ``Model``.  The builders restate the model definitions of the
reference drivers:

* ``hm2x2``   drivers/cdn_hm_2dsquare.f90:46,96-108,221-259 (BASELINE configs K1-K3)
* ``bhz2``    drivers/cdn_bhz_2d.f90:47-49,213-248,290-299   (K4, complex H)
* ``hm_ns18`` 3x2 open Hubbard cluster, Nbath=2              (K5)

Arrays are numpy, Fortran (column-major) ordered, exactly the memory layout the
C ABI (include/cdmft_b200.h) expects.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Model:
    nlat: int
    norb: int
    nspin: int
    nbath: int
    imphloc: np.ndarray  # complex128 [Nlat,Nlat,Nspin,Nspin,Norb,Norb], order='F'
    hbath: np.ndarray  # complex128 [Nlat,Nlat,Nspin,Nspin,Norb,Norb,Nbath], order='F'
    vbath: np.ndarray  # float64 [Nlso,Nbath], order='F'; Nlso index = index_stride_lso
    uloc: np.ndarray = field(default_factory=lambda: np.array([2.0, 0, 0, 0, 0]))
    ust: float = 0.0
    jh: float = 0.0
    jx: float = 0.0
    jp: float = 0.0
    xmu: float = 0.0
    hfmode: bool = True
    name: str = "model"

    def __post_init__(self):
        L, S, O, B = self.nlat, self.nspin, self.norb, self.nbath
        self.imphloc = np.asfortranarray(self.imphloc, dtype=np.complex128)
        self.hbath = np.asfortranarray(self.hbath, dtype=np.complex128)
        self.vbath = np.asfortranarray(self.vbath, dtype=np.float64)
        u = np.zeros(5)
        u[: len(np.atleast_1d(self.uloc))] = np.atleast_1d(self.uloc)
        self.uloc = u
        assert self.imphloc.shape == (L, L, S, S, O, O)
        assert self.hbath.shape == (L, L, S, S, O, O, B)
        assert self.vbath.shape == (L * S * O, B)

    @property
    def nimp(self) -> int:
        return self.nlat * self.norb

    @property
    def ns(self) -> int:
        # ed_setup_dimensions, ED_SETUP.f90:111-120
        return self.nimp * (self.nbath + 1)

    @property
    def is_real(self) -> bool:
        return bool(np.all(self.imphloc.imag == 0) and np.all(self.hbath.imag == 0))

    @property
    def jhflag(self) -> bool:
        # ED_SETUP.f90:200-201
        return self.norb > 1 and (self.jx != 0.0 or self.jp != 0.0)


def _open_cluster_hloc(nx: int, ny: int, ts: float) -> np.ndarray:
    """-ts on the nearest-neighbour bonds of an open nx*ny cluster, site index
    Nx*(y-1)+x (drivers/cdn_hm_2dsquare.f90:221-259,307-313)."""
    n = nx * ny
    h = np.zeros((n, n), dtype=np.complex128)
    for x in range(nx):
        for y in range(ny):
            a = nx * y + x
            if x + 1 < nx:
                b = nx * y + x + 1
                h[a, b] = h[b, a] = -ts
            if y + 1 < ny:
                b = nx * (y + 1) + x
                h[a, b] = h[b, a] = -ts
    return h


def hubbard_cluster(nx: int, ny: int, nbath: int, ts: float = 0.25, uloc: float = 2.0,
                    hwband: float = 2.0, name: str | None = None) -> Model:
    """Norb=1, Nspin=1 Hubbard cluster with a replica bath: Hsym1 = 1, Hsym2 = |Hloc|,
    lambda(ib,1)=onsite_ib, lambda(ib,2)=1, V = max(0.1, 1/sqrt(Nbath))
    (drivers/cdn_hm_2dsquare.f90:96-108; ED_BATH/dmft_aux.f90:70)."""
    nlat = nx * ny
    hl = _open_cluster_hloc(nx, ny, ts)
    imphloc = np.zeros((nlat, nlat, 1, 1, 1, 1), dtype=np.complex128, order="F")
    imphloc[:, :, 0, 0, 0, 0] = hl
    onsite = np.zeros(nbath)
    if nbath > 1:
        for ib in range(nbath):
            onsite[ib] = (ib - (nbath - 1) / 2.0) * 2 * hwband / (nbath - 1)
        if nbath % 2 == 0:
            onsite[nbath // 2 - 1] = -0.1
            onsite[nbath // 2] = 0.1
    # Nbath=1 divides by zero in the driver (:101) -> onsite = 0
    hbath = np.zeros((nlat, nlat, 1, 1, 1, 1, nbath), dtype=np.complex128, order="F")
    for ib in range(nbath):
        hbath[:, :, 0, 0, 0, 0, ib] = onsite[ib] * np.eye(nlat) + 1.0 * np.abs(hl)
    v = max(0.1, 1.0 / np.sqrt(nbath))
    vbath = np.full((nlat, nbath), v, order="F")
    return Model(nlat, 1, 1, nbath, imphloc, hbath, vbath, uloc=np.array([uloc]),
                 name=name or f"hm{nx}x{ny}_nbath{nbath}")


def hm2x2(nbath: int) -> Model:
    """BASELINE configs K1 (Nbath=1, Ns=8), K2 (Nbath=2, Ns=12), K3 (Nbath=3, Ns=16)."""
    return hubbard_cluster(2, 2, nbath, name=f"hm2x2_nbath{nbath}")


def hm_ns18() -> Model:
    """BASELINE config K5: 3x2 open cluster, Nbath=2 -> Ns=18."""
    return hubbard_cluster(3, 2, 2, name="hm3x2_nbath2_ns18")


def bhz2(nbath: int = 3, kanamori: bool = False, mh: float = 1.0, ts: float = 0.25,
         lam: float = 0.3, uloc: float = 2.0, hwband: float = 2.0) -> Model:
    """BASELINE config K4: 2-site BHZ cluster, Norb=2, complex hoppings
    (drivers/cdn_bhz_2d.f90:213-248,290-299): on-site Mh*sigma_z, x-hop
    T_x = -ts*sigma_z + (i/2)*lam*sigma_x at hopping(ind2,ind1) and T_x^dagger the other way.
    Bath replicas: lambda_sym = (Mh*r_ib, ts, lam), r = linspace(hwband/Nbath, hwband, Nbath)."""
    nlat, norb = 2, 2
    sz = np.diag([1.0, -1.0]).astype(np.complex128)
    sx = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    tx = -ts * sz + 0.5j * lam * sx

    def build(m, t, l):
        h = np.zeros((nlat, nlat, 1, 1, norb, norb), dtype=np.complex128, order="F")
        txx = -t * sz + 0.5j * l * sx
        for a in range(nlat):
            h[a, a, 0, 0] = m * sz
        h[1, 0, 0, 0] = txx
        h[0, 1, 0, 0] = txx.conj().T
        return h

    del tx
    imphloc = build(mh, ts, lam)
    hbath = np.zeros((nlat, nlat, 1, 1, norb, norb, nbath), dtype=np.complex128, order="F")
    r = np.linspace(hwband / nbath, hwband, nbath)
    for ib in range(nbath):
        hbath[..., ib] = build(mh * r[ib], ts, lam)
    vbath = np.full((nlat * norb, nbath), 1.0 / np.sqrt(nbath), order="F")
    kw = {}
    if kanamori:
        jh = 0.25 * uloc
        kw = dict(jh=jh, jx=jh, jp=jh, ust=uloc - 2 * jh)
    return Model(nlat, norb, 1, nbath, imphloc, hbath, vbath, uloc=np.array([uloc, uloc]),
                 name=f"bhz2_nbath{nbath}" + ("_kanamori" if kanamori else ""), **kw)


def random_model(nlat: int, norb: int, nbath: int, nspin: int = 1, complex_h: bool = True,
                 seed: int = 0, kanamori: bool = False, hfmode: bool = True) -> Model:
    """Random Hermitian one-body matrices; exercises every branch of the term lists
    (off-diagonal replica hoppings, spin-dependent blocks when nspin=2, complex phases)."""
    rng = np.random.default_rng(seed)
    n = nlat * norb

    def herm(scale):
        a = rng.normal(size=(n, n)) + (1j * rng.normal(size=(n, n)) if complex_h else 0)
        a = scale * (a + a.conj().T) / 2
        a[np.abs(a) < 0.35 * scale] = 0  # make it sparse-ish, keep hermiticity
        a = (a + a.conj().T) / 2
        return a

    def to6(mats):  # list over spin of [n,n] (index = iorb + ilat*norb) -> [L,L,S,S,O,O]
        h = np.zeros((nlat, nlat, nspin, nspin, norb, norb), dtype=np.complex128, order="F")
        for s, m in enumerate(mats):
            for a in range(n):
                for b in range(n):
                    h[a // norb, b // norb, s, s, a % norb, b % norb] = m[a, b]
        return h

    imphloc = to6([herm(0.5) for _ in range(nspin)])
    hbath = np.zeros((nlat, nlat, nspin, nspin, norb, norb, nbath), dtype=np.complex128, order="F")
    for ib in range(nbath):
        hbath[..., ib] = to6([herm(0.7) + np.diag(rng.normal(size=n)) for _ in range(nspin)])
    vbath = np.asfortranarray(rng.normal(size=(nlat * nspin * norb, nbath)))
    vbath[np.abs(vbath) < 0.2] = 0.0
    kw = {}
    uloc = np.zeros(5)
    uloc[:norb] = 1.0 + rng.random(norb)
    if norb > 1:
        kw = dict(ust=0.8, jh=0.3)
        if kanamori:
            kw.update(jx=0.25, jp=0.2)
    return Model(nlat, norb, nspin, nbath, imphloc, hbath, vbath, uloc=uloc, xmu=0.3,
                 hfmode=hfmode, name=f"random_L{nlat}O{norb}B{nbath}S{nspin}", **kw)


def get_sector(ns: int, nup: int, ndw: int) -> int:
    """isector = 1 + Nup*(Ns+1) + Ndw (get_Sector, ED_SETUP.f90:446-457)."""
    return 1 + nup * (ns + 1) + ndw
