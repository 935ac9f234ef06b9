"""Generates VKPM_TABLE_RC / VKPM_TABLE_L2 of vulkpy_b200/csrc/vkp_math.cuh.

log2(x), x = 2^e m, m in [2/3, 4/3): interval i = 5 bits of (bits(x) - bits(2/3)) >> 18.  rc_i ~ 1 / centre_i is
rounded to 21 significant bits so that (a) as a binary64 number its low word is zero -- the device keeps only the
high word per lane and a lookup is ONE warp shuffle, no float -> double conversion -- and (b) r = m rc_i - 1 is exact
in binary64 (24 + 21 bits).  l2_i = -log2(rc_i), correctly rounded (mpmath, 120 bits).  Interval 21 contains 1.0 and
gets rc = 1, l2 = 0 exactly, hence log2(1) = 0 and log2(2^n) = n.
"""
import struct
import mpmath as mp

mp.mp.prec = 160
BASE = 0x3f2aaaab


def f32(bits):
    return mp.mpf(struct.unpack("<f", struct.pack("<I", bits))[0])


def round_bits(x, nbits):
    m, e = mp.frexp(x)              # x = m 2^e, 0.5 <= m < 1
    q = mp.floor(m * 2 ** nbits + mp.mpf(0.5))
    return q * mp.mpf(2) ** (e - nbits)


rc, l2 = [], []
worst = 0
for i in range(32):
    lo, hi = f32(BASE + (i << 18)), f32(BASE + ((i + 1) << 18) - 1)
    c = 2 / (lo + hi)
    r = mp.mpf(1) if lo <= 1 <= hi else round_bits(c, 21)
    rc.append(r)
    l2.append(-mp.log(r, 2))
    worst = max(worst, abs(lo * r - 1), abs(hi * r - 1))
print("// max |r| =", mp.nstr(worst, 6))
print("#define VKPM_TABLE_RC \\")
for k in range(0, 32, 4):
    print("  " + ", ".join(float(x).hex() for x in rc[k:k + 4]) + ", \\")
print("\n#define VKPM_TABLE_L2 \\")
for k in range(0, 32, 3):
    print("  " + ", ".join(float(x).hex() for x in l2[k:k + 3]) + ", \\")
for x in rc:       # 21 significant bits: the low 32 bits of the double are zero
    assert struct.unpack("<Q", struct.pack("<d", float(x)))[0] & 0xffffffff == 0
