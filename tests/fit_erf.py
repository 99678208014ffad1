"""Fit and check the single-branch erf used by the trunk kernel's GELU epilogue (st_gemm_tc.cu: gelu_sb):
    erf(z) = 1 - 2^(-z q(z)),  q = degree-10 polynomial fitted (Chebyshev nodes, float64) to -log2(erfc(z)) / z on [0, 4.2],
    z clamped to 4.2 (erf(4.2) rounds to 1 in fp32).  python tests/fit_erf.py  prints the coefficients and the error of the
    fp32 evaluation; tests/test_packer_cpu.py::test_gelu_single_branch_erf pins the table that is compiled into the kernel."""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import erf, erfc

X, DEG = 4.2, 10
COEF = [1.62790707e+00, 9.18446271e-01, 1.48284197e-01, -2.76306103e-02, -2.98958275e-04, 2.56001420e-03,
        -1.07292213e-03, 2.56761395e-04, -3.83041166e-05, 3.31221616e-06, -1.26982216e-07]          # ascending powers of z


def fit():
    xs = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * X / 2 + X / 2
    c = C.Chebyshev.fit(xs, -np.log2(erfc(xs)) / xs, DEG, domain=[0, X])
    return c.convert(kind=np.polynomial.Polynomial).coef


def gelu_f32(x, coef=COEF):
    """The kernel's arithmetic in numpy float32 (Horner with separate multiply / add roundings is the pessimistic case)."""
    x = x.astype(np.float32)
    z = np.minimum(np.abs(x) * np.float32(0.70710678118654752440), np.float32(X)).astype(np.float32)
    acc = np.full_like(z, np.float32(coef[-1]))
    for k in range(len(coef) - 2, -1, -1):
        acc = (acc * z + np.float32(coef[k])).astype(np.float32)
    e = np.exp2((-(z * acc)).astype(np.float32)).astype(np.float32)
    er = np.copysign((np.float32(1) - e).astype(np.float32), x)
    h = (x * np.float32(0.5)).astype(np.float32)
    return (h * er + h).astype(np.float32)


def errors():
    x = np.linspace(-10, 10, 2000001)
    ref = x * 0.5 * (1 + erf(x / np.sqrt(2)))
    got = gelu_f32(x).astype(np.float64)
    ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    return float(np.abs(got - ref).max()), float((np.abs(got - ref) / np.maximum(ulp, 1e-12))[np.abs(ref) > 1e-3].max())


if __name__ == "__main__":
    print("fitted:", ", ".join(f"{v:.8e}" for v in fit()))
    a, u = errors()
    print(f"GELU fp32 evaluation: max abs error {a:.3e}, max error in ulps of the result (|gelu| > 1e-3) {u:.2f}")
