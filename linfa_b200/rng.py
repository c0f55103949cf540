"""Host RNGs for the binding.  `Xoshiro256Plus` is rand_xoshiro 0.6's generator,
the default rng of `KMeans::params` (KM/algorithm.rs:211: seed_from_u64(42))."""

_M = 0xFFFFFFFFFFFFFFFF


class Xoshiro256Plus:
    __slots__ = ("s",)

    def __init__(self, state):
        self.s = [int(v) & _M for v in state]

    @classmethod
    def seed_from_u64(cls, seed):
        x = int(seed) & _M
        out = []
        for _ in range(4):  # SplitMix64 fills the state (rand_core SeedableRng::seed_from_u64)
            x = (x + 0x9E3779B97F4A7C15) & _M
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
            out.append(z ^ (z >> 31))
        if not any(out):
            return cls.seed_from_u64(0)
        return cls(out)

    def clone(self):
        return Xoshiro256Plus(self.s)

    def next_u64(self):
        s = self.s
        r = (s[0] + s[3]) & _M
        t = (s[1] << 17) & _M
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = ((s[3] << 45) | (s[3] >> 19)) & _M
        return r

    def next_u32(self):
        return self.next_u64() >> 32
