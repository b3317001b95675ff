"""ORACLE (test infrastructure only). Duplex-sponge Fiat-Shamir transcript, restating
challenger/challenger.go of the reference."""
from . import goldilocks as gl
from .poseidon import GoldilocksChip, BN254Chip, SPONGE_RATE, SPONGE_WIDTH


class FriChallenges:
    pass


class Chip:
    def __init__(self, api):
        self.api = api
        self.poseidonChip = GoldilocksChip(api)
        self.poseidonBN254Chip = BN254Chip(api)
        self.glApi = gl.Chip(api)
        self.spongeState = [0] * SPONGE_WIDTH
        self.inputBuffer = []
        self.outputBuffer = []
        self.n_duplex = 0

    def ObserveElement(self, e):
        """challenger.go:42-49"""
        self.outputBuffer = []
        self.inputBuffer.append(e)
        if len(self.inputBuffer) == SPONGE_RATE:
            self.duplexing()

    def ObserveElements(self, es):
        for e in es:
            self.ObserveElement(e)

    def ObserveHash(self, h):
        self.ObserveElements(list(h))

    def ObserveBN254Hash(self, h):
        self.ObserveElements(self.poseidonBN254Chip.ToVec(h))

    def ObserveCap(self, cap):
        for h in cap:
            self.ObserveBN254Hash(h)

    def ObserveExtensionElement(self, e):
        self.ObserveElements(list(e))

    def ObserveExtensionElements(self, es):
        for e in es:
            self.ObserveExtensionElement(e)

    def ObserveOpenings(self, openings):
        for batch in openings:
            self.ObserveExtensionElements(batch)

    def GetChallenge(self):
        """challenger.go:89-98 (pops from the END)"""
        if len(self.inputBuffer) != 0 or len(self.outputBuffer) == 0:
            self.duplexing()
        return self.outputBuffer.pop()

    def GetNChallenges(self, n):
        return [self.GetChallenge() for _ in range(n)]

    def GetExtensionChallenge(self):
        v = self.GetNChallenges(2)
        return (v[0], v[1])

    def GetHash(self):
        return [self.GetChallenge() for _ in range(4)]

    def GetFriChallenges(self, commit_phase_merkle_caps, final_poly, pow_witness, config):
        """challenger.go:117-144"""
        c = FriChallenges()
        c.FriAlpha = self.GetExtensionChallenge()
        c.FriBetas = []
        for cap in commit_phase_merkle_caps:
            self.ObserveCap(cap)
            c.FriBetas.append(self.GetExtensionChallenge())
        self.ObserveExtensionElements(final_poly)
        self.ObserveElement(pow_witness)
        c.FriPowResponse = self.GetChallenge()
        c.FriQueryIndices = self.GetNChallenges(config.NumQueryRounds)
        return c

    def duplexing(self):
        """challenger.go:146-166"""
        assert len(self.inputBuffer) <= SPONGE_RATE
        self.n_duplex += 1
        for i, v in enumerate(self.inputBuffer):
            self.spongeState[i] = self.glApi.Reduce(v)
        self.inputBuffer = []
        self.spongeState = self.poseidonChip.Poseidon(self.spongeState)
        self.outputBuffer = list(self.spongeState[:SPONGE_RATE])
