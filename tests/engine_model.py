"""numpy model of the device FFT engine's index math (test infrastructure, not product).

Mirrors fourierconvolutioncudalib_b200/csrc/fft_engine.cuh: in-place mixed-radix DIF forward
(result in digit-reversed *positions*), mirrored DIT inverse, the even-length R2C split /
C2R merge done on the packed half-length transform, and the shared-memory swizzle used by
the transposing loads of the X pass.  Used by tests/test_engine_model.py to pin the math on
the CPU before the same formulas run on the GPU.
"""
import numpy as np

FAST_RADICES = (2, 3, 4, 5, 7, 8, 16)

_POW2_PLAN = {0: [], 1: [2], 2: [4], 3: [8], 4: [16], 5: [8, 4], 6: [8, 8], 7: [16, 8], 8: [16, 16], 9: [8, 8, 8],
              10: [16, 16, 4], 11: [16, 16, 8], 12: [16, 16, 16]}


def factorize(L, style=0):
    """Same result as the C planner (fc_plan.cu: factorize): fewest stages for the power-of-two part
    with radices <= 16, then the odd 7-smooth part (composite radices 15 / 9 / 6 / 12 / 10 where factors pair up), then
    the remaining primes (generic stages)."""
    # two stages of fat composite radices for the measured lengths (bit 0: y axis, bit 1: fused z axis, bit 2: x axis)
    for fl, styles, r0, r1 in ((300, 3, 20, 15), (420, 3, 20, 21), (270, 3, 18, 15), (448, 1, 16, 28), (560, 2, 28, 20),
                               (150, 4, 10, 15), (135, 4, 9, 15)):
        if fl == L and (styles >> style) & 1:
            return [r0, r1]
    out = []
    n = L
    e = 0
    while n > 1 and n % 2 == 0:
        n //= 2
        e += 1
    while e > 12:
        out.append(16)
        e -= 4
    if style == 1 and L == 256:
        out += [8, 8, 4]
    elif style == 2 and L == 1024:
        out += [16, 8, 8]
    elif style == 2 and L == 512:
        out += [16, 4, 8]
    else:
        out += _POW2_PLAN[e]
    # odd 7-smooth part: composite register stages (15, 9) and a leftover 3 / 5 joining a trailing 2 / 4
    cnt = {}
    for r in (3, 5, 7):
        cnt[r] = 0
        while n % r == 0 and n > 1:
            cnt[r] += 1
            n //= r
    comp = []
    if len(out) + cnt[3] + cnt[5] + cnt[7] > (4 if style == 2 else 3):   # x axis: up to four stages of primes; y / z: three
        while cnt[3] >= 1 and cnt[5] >= 1:
            comp.append(15)
            cnt[3] -= 1
            cnt[5] -= 1
        if cnt[3] % 2 == 1 and out and out[-1] in (2, 4):
            out[-1] *= 3
            cnt[3] -= 1
        while cnt[3] >= 2:
            comp.append(9)
            cnt[3] -= 2
        if cnt[5] >= 1 and out and out[-1] == 2:
            out[-1] = 10
            cnt[5] -= 1
    out += comp + [3] * cnt[3] + [5] * cnt[5] + [7] * cnt[7]
    p = 11
    while n > 1:
        while n % p == 0:
            out.append(p)
            n //= p
        p += 2
    if not out:
        out = [1]
    return out


def twiddles(L):
    t = np.arange(L)
    return np.exp(-2j * np.pi * t / L)


def rev_positions(L, radices):
    """freq index held at position p after the forward in-place DIF."""
    rev = np.zeros(L, dtype=np.int64)
    for p in range(L):
        k, mul, rem, Li = 0, 1, p, L
        for R in radices:
            S = Li // R
            m, rem = divmod(rem, S)
            k += m * mul
            mul *= R
            Li = S
        rev[p] = k
    return rev


def fwd_inplace(x, radices):
    x = np.array(x, dtype=np.complex128)
    L = x.shape[0]
    tw = twiddles(L)
    Li = L
    for R in radices:
        S = Li // R
        step = L // Li
        for b in range(L // R):
            beta, j = divmod(b, S)
            base = beta * Li + j
            idx = base + S * np.arange(R)
            v = x[idx]
            y = np.array([sum(v[k] * tw[((k * m) % R) * (L // R)] for k in range(R)) for m in range(R)])
            y = y * tw[(j * np.arange(R) * step)]
            x[idx] = y
        Li = S
    return x


def inv_inplace(x, radices):
    x = np.array(x, dtype=np.complex128)
    L = x.shape[0]
    tw = twiddles(L)
    Li = 1
    for R in reversed(radices):
        S = Li
        Li = S * R
        step = L // Li
        for b in range(L // R):
            beta, j = divmod(b, S)
            base = beta * Li + j
            idx = base + S * np.arange(R)
            v = x[idx] * np.conj(tw[(j * np.arange(R) * step)])
            y = np.array([sum(v[k] * np.conj(tw[((k * m) % R) * (L // R)]) for k in range(R)) for m in range(R)])
            x[idx] = y
    return x


def r2c_even(xreal, radices_half):
    """Packed R2C of even length n=2M: returns X[0..M] in *position* order (pos M = Nyquist)."""
    n = xreal.shape[0]
    M = n // 2
    z = xreal[0::2] + 1j * xreal[1::2]
    Zp = fwd_inplace(z, radices_half)
    rev = rev_positions(M, radices_half)
    P = np.zeros(M, dtype=np.int64)
    P[rev] = np.arange(M)
    wn = np.exp(-2j * np.pi * np.arange(M + 1) / n)
    out = np.zeros(M + 1, dtype=np.complex128)
    Z0 = Zp[P[0]]
    out[P[0]] = Z0.real + Z0.imag
    out[M] = Z0.real - Z0.imag
    for k in range(1, M // 2 + 1):
        k2 = M - k
        a, b = Zp[P[k]], Zp[P[k2]]
        E = 0.5 * (a + np.conj(b))
        O = -0.5j * (a - np.conj(b))
        Xk = E + wn[k] * O
        Xk2 = np.conj(E - wn[k] * O)
        out[P[k]] = Xk
        out[P[k2]] = Xk2
    return out, P


def c2r_even(Xpos, n, radices_half):
    M = n // 2
    rev = rev_positions(M, radices_half)
    P = np.zeros(M, dtype=np.int64)
    P[rev] = np.arange(M)
    wn = np.exp(-2j * np.pi * np.arange(M + 1) / n)
    Z = np.zeros(M, dtype=np.complex128)
    X0, XM = Xpos[P[0]], Xpos[M]
    Z[P[0]] = (X0.real + XM.real) + 1j * (X0.real - XM.real)
    for k in range(1, M // 2 + 1):
        k2 = M - k
        a, b = Xpos[P[k]], Xpos[P[k2]]
        s = a + np.conj(b)
        d = (a - np.conj(b)) * np.conj(wn[k])
        Z[P[k]] = s + 1j * d
        Z[P[k2]] = np.conj(s) + 1j * np.conj(d)   # derived from the k -> M-k symmetry
    z = inv_inplace(Z, radices_half)
    out = np.zeros(n)
    out[0::2] = z.real
    out[1::2] = z.imag
    return out


def swz(p):
    return (p ^ (p >> 3)) & 7
