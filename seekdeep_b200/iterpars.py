"""Par-file iteration tables (njhseq CollapseIterations / IterPar) as qluster consumes them:
format documented in clusterDownSetUp.cpp:63-84, files in etc/SeekDeepParameterFiles/*,
technology defaults chosen in clusterDownSetUp.cpp:241-312."""
from dataclasses import dataclass
from typing import List


@dataclass
class IterPar:
    stopCheck: int
    smallCutoff: int
    oneBaseIndel: float = 0.0
    twoBaseIndel: float = 0.0
    largeBaseIndel: float = 0.0
    hqMismatches: float = 0.0
    lqMismatches: float = 0.0
    lowKmerMismatches: float = 0.0
    onPerId: bool = False
    onHqPerId: bool = False
    perId: float = 0.0


class CollapseIterations:
    HEADER = "stopCheck:smallCutoff:1baseIndel:2baseIndel:>2baseIndel:HQMismatches:LQMismatches:LKMismatches"

    def __init__(self, iters: List[IterPar]):
        self.iters = list(iters)

    def __len__(self):
        return len(self.iters)

    def __iter__(self):
        return iter(self.iters)

    @classmethod
    def parse(cls, text, onPerId=False, onHqPerId=False):
        """processIteratorMap / processIteratorMapOnPerId (clusterDownSetUp.cpp:415-438)."""
        iters = []
        for raw in text.splitlines():
            line = raw.strip()
            if not line or line.startswith("#"):
                continue
            tok = line.split(":")
            if not tok[0].replace(".", "", 1).isdigit():
                continue  # header line
            if onPerId:
                if len(tok) != 3:
                    raise ValueError(f"per-id par line needs 3 columns: {line!r}")
                iters.append(IterPar(int(tok[0]), int(tok[1]), onPerId=True, onHqPerId=onHqPerId,
                                     perId=float(tok[2])))
            else:
                if len(tok) != 8:
                    raise ValueError(f"par line needs 8 columns: {line!r}")
                iters.append(IterPar(int(tok[0]), int(tok[1]), *(float(x) for x in tok[2:])))
        if not iters:
            raise ValueError("no iterations in par file")
        return cls(iters)

    @classmethod
    def from_file(cls, path, **kw):
        with open(path) as fh:
            return cls.parse(fh.read(), **kw)

    @classmethod
    def gen_illumina_default_pars(cls, stopCheck=100):
        """CollapseIterations::genIlluminaDefaultPars(checkAmount) (clusterDownSetUp.cpp:266-268);
        same table as etc/SeekDeepParameterFiles/illumina_lkmer1:2-23 with stopCheck substituted."""
        lq = [0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 8]
        one = [1] + [2] * 10
        its = []
        for small in (3, 0):
            for o, l in zip(one, lq):
                its.append(IterPar(stopCheck, small, o, 0, 0, 0, l, 1))
        return cls(its)

    @classmethod
    def gen_otu_pars(cls, stopCheck, perId, onHqPerId=True):
        """CollapseIterations::genOtuPars (clusterDownSetUp.cpp:303-305); etc/.../otu_99:2-5."""
        return cls([IterPar(stopCheck, s, onPerId=True, onHqPerId=onHqPerId, perId=perId) for s in (3, 3, 0, 0)])

    def write_pars(self):
        out = [self.HEADER]
        for it in self.iters:
            vals = [it.stopCheck, it.smallCutoff, it.oneBaseIndel, it.twoBaseIndel, it.largeBaseIndel,
                    it.hqMismatches, it.lqMismatches, it.lowKmerMismatches]
            out.append(":".join(f"{v:g}" for v in vals))
        return "\n".join(out) + "\n"
