"""Named parameter sets (values from the reference's hard-coded constants and SURVEY.md 8(d))."""
from __future__ import annotations

from dataclasses import asdict, dataclass

from . import abi


@dataclass(frozen=True)
class Params:
    n: int          # LWE dimension
    N: int          # ring degree
    k: int          # TRLWE mask polynomials
    l: int          # bootstrapping-key gadget levels
    Bg_bit: int     # log2 gadget base
    t: int          # key-switch levels
    base_bit: int   # log2 key-switch base
    lwe_sigma: float = 2.0 ** -15
    rlwe_sigma: float = 2.0 ** -44

    def c(self) -> abi.ParamsS:
        return abi.ParamsS(self.n, self.N, self.k, self.l, self.Bg_bit, self.t, self.base_bit)

    @property
    def bsk_bytes(self) -> int:
        return self.n * (self.k + 1) * self.l * (self.k + 1) * self.N * 8

    @property
    def ksk_bytes(self) -> int:
        return self.k * self.N * self.t * ((1 << self.base_bit) - 1) * (self.n + 1) * 8

    def flops_per_step(self) -> int:
        """(T_f + T_i)(5 M log2 M + 6 M) + 8 M (k+1) T_f, M = N/2."""
        M = self.N // 2
        lg = M.bit_length() - 1
        Tf, Ti = (self.k + 1) * self.l, self.k + 1
        return (Tf + Ti) * (5 * M * lg + 6 * M) + 8 * M * (self.k + 1) * Tf

    def flops_per_pbs(self) -> int:
        return self.n * self.flops_per_step()

    def asdict(self):
        return asdict(self)


# test/benchmark.c:66-75 (TFHEpp Level 2): BASELINE configs[0] and configs[2]
LEVEL2 = Params(n=632, N=2048, k=1, l=4, Bg_bit=9, t=8, base_bit=4, lwe_sigma=2.0 ** -15, rlwe_sigma=2.0 ** -44)
# "TFHEpp-Level-1-style" in the 64-bit torus: BASELINE configs[1]
LEVEL1 = Params(n=632, N=1024, k=1, l=3, Bg_bit=6, t=7, base_bit=2, lwe_sigma=2.0 ** -15, rlwe_sigma=2.0 ** -25)
# test/benchmark.c:53-54 SET_1
SET_1 = Params(n=585, N=1024, k=1, l=2, Bg_bit=8, t=5, base_bit=2, lwe_sigma=9.141776004202573e-5,
               rlwe_sigma=2.989040792967434e-8)
# test/tests.c:44-45 SET_2 (the reference unit tests' default)
SET_2 = Params(n=744, N=2048, k=1, l=1, Bg_bit=23, t=5, base_bit=3, lwe_sigma=7.747831515176779e-6,
               rlwe_sigma=2.2148688116005568e-16)

# circuit-bootstrap shape of scripts/bench_next.py config 4b: Level-2 gadget on the N = 1024 ring
CB_1024 = Params(n=632, N=1024, k=1, l=4, Bg_bit=9, t=6, base_bit=4, lwe_sigma=2.0 ** -30, rlwe_sigma=2.0 ** -55)

NAMED = {"level2": LEVEL2, "level1": LEVEL1, "set1": SET_1, "set2": SET_2, "cb1024": CB_1024}
