"""Near-minimax polynomial coefficients for the trimmed FP64 routines of lf_math.cuh (math v4 - v6).
Remez exchange in 40-digit arithmetic (mpmath) on the scaled variable u = x / h; prints the coefficients and the
achieved maximum absolute error.    python tools/math/fit_coeffs.py"""
import mpmath as mp

mp.mp.dps = 40


def remez(f, powers, h, iters=12, ngrid=4001):
    """minimax fit of f(x) ~ sum_k c_k x^p_k on [-h, h] (absolute error); returns (c, max_err)."""
    m = len(powers)
    u = [-mp.cos(mp.pi * i / m) for i in range(m + 1)]
    grid = [mp.mpf(-1) + mp.mpf(2) * i / (ngrid - 1) for i in range(ngrid)]
    fg = [f(h * g) for g in grid]
    c = None
    for _ in range(iters):
        A = mp.matrix(m + 1, m + 1)
        b = mp.matrix(m + 1, 1)
        for i, ui in enumerate(u):
            for k, p in enumerate(powers):
                A[i, k] = ui ** p
            A[i, m] = (-1) ** i
            b[i] = f(h * ui)
        sol = mp.lu_solve(A, b)
        c = [sol[k] for k in range(m)]
        err = [sum(c[k] * g ** p for k, p in enumerate(powers)) - fg[i] for i, g in enumerate(grid)]
        ext = [0] + [i for i in range(1, ngrid - 1) if (err[i] - err[i - 1]) * (err[i + 1] - err[i]) <= 0] + [ngrid - 1]
        ext = sorted(sorted(ext, key=lambda i: -abs(err[i]))[:m + 1])
        if len(ext) == m + 1:
            u = [grid[i] for i in ext]
    emax = max(abs(e) for e in err)
    return [c[k] / h ** p for k, p in enumerate(powers)], emax


def remez_ab(f, deg, a, b, iters=14, ngrid=4001):
    """minimax fit of f(x) ~ sum_k c_k x^k on [a, b] (absolute error), coefficients in x itself; returns (c, max_err)."""
    m = deg + 1
    mid, h = (a + b) / 2, (b - a) / 2
    u = [-mp.cos(mp.pi * i / m) for i in range(m + 1)]
    grid = [mp.mpf(-1) + mp.mpf(2) * i / (ngrid - 1) for i in range(ngrid)]
    fg = [f(mid + h * g) for g in grid]
    c, err = None, None
    for _ in range(iters):
        A = mp.matrix(m + 1, m + 1)
        b_ = mp.matrix(m + 1, 1)
        for i, ui in enumerate(u):
            x = mid + h * ui
            for k in range(m):
                A[i, k] = x ** k
            A[i, m] = (-1) ** i
            b_[i] = f(x)
        sol = mp.lu_solve(A, b_)
        c = [sol[k] for k in range(m)]
        err = [sum(c[k] * (mid + h * g) ** k for k in range(m)) - fg[i] for i, g in enumerate(grid)]
        ext = [0] + [i for i in range(1, ngrid - 1) if (err[i] - err[i - 1]) * (err[i + 1] - err[i]) <= 0] + [ngrid - 1]
        ext = sorted(sorted(ext, key=lambda i: -abs(err[i]))[:m + 1])
        if len(ext) == m + 1:
            u = [grid[i] for i in ext]
    return c, max(abs(e) for e in err)


if __name__ == '__main__':
    h = mp.mpf(2) ** -9
    for name, f in (("log1p(x)", mp.log1p), ("2^x", lambda x: mp.power(2, x))):
        for deg in (3, 4):
            c, e = remez(f, list(range(0, deg + 1)), h)
            print("%s degree %d on |x| <= 2^-9: max abs err %s" % (name, deg, mp.nstr(e, 3)))
            print("    " + ", ".join(mp.nstr(v, 20) for v in c))
            if name == "2^x" and deg == 3:
                # math v6: C0 (1 + c1 x + c2 x^2 + c3 x^3) -- C0 goes into the table entries (EXP_TAB_SCALE), the last Horner
                # step adds the immediate 1.0, c3 is rounded to its high word
                print("    normalised by the constant term: " + ", ".join(mp.nstr(v / c[0], 22) for v in c))
    # math v6: log1p with the linear coefficient FIXED at 1 (the last-but-one Horner step adds the immediate 1.0)
    c, e = remez(lambda x: mp.log1p(x) - x, [0, 2, 3], h)
    print("log1p(x) - x ~ c0 + c2 x^2 + c3 x^3 on |x| <= 2^-9: max abs err %s" % mp.nstr(e, 3))
    print("    " + ", ".join(mp.nstr(v, 22) for v in c))
    # (git history: the decay factor with a masked-index look-up, remainder in [-2^-12, 15 2^-12) -- LF_EXP_MASKED, measured
    #  and dropped in round 2)
    c, e = remez_ab(lambda x: mp.power(2, x), 3, -mp.mpf(2) ** -12, 15 * mp.mpf(2) ** -12)
    print("2^x degree 3 on [-2^-12, 15 * 2^-12]: max abs err %s" % mp.nstr(e, 3))
    print("    " + ", ".join(mp.nstr(v, 20) for v in c))
