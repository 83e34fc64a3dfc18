"""FASTQ ingest (clusterDown.cpp:207-232): plain and gzip files, CRLF, a last record without a newline,
lower-case bases, empty files; malformed input is refused with the record number.  Host code only: no GPU needed."""
import gzip

import numpy as np
import pytest

import seekdeep_b200 as sd
from seekdeep_b200.qluster import read_fastq


def _fastq(records, nl="\n", last_newline=True):
    txt = nl.join(f"@{n}{nl}{s}{nl}+{nl}{q}" for n, s, q in records)
    return txt + (nl if last_newline and records else "")


def _records(rng, n):
    recs = []
    for i in range(n):
        ln = int(rng.integers(1, 400))
        s = "".join(rng.choice(list("ACGTN"), ln))
        q = "".join(chr(33 + int(x)) for x in rng.integers(0, 42, ln))
        recs.append((f"read{i} extra:field", s, q))
    return recs


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("nl,last_newline", [("\n", True), ("\r\n", True), ("\n", False)])
def test_read_fastq_round_trip(tmp_path, gz, nl, last_newline):
    rng = np.random.default_rng(4)
    recs = _records(rng, 3000)  # > 1 MiB of text: several reader buffers
    path = tmp_path / ("reads.fastq.gz" if gz else "reads.fastq")
    data = _fastq(recs, nl, last_newline).encode()
    if gz:
        with gzip.open(path, "wb") as f:
            f.write(data)
    else:
        path.write_bytes(data)
    bases, quals, offsets, names = read_fastq(path)
    assert len(names) == len(recs) and names == [r[0] for r in recs]
    assert bases.tobytes().decode() == "".join(r[1] for r in recs)
    assert np.array_equal(quals, np.frombuffer("".join(r[2] for r in recs).encode(), np.uint8) - 33)
    assert np.array_equal(np.diff(offsets), [len(r[1]) for r in recs])


def test_read_fastq_edge_cases(tmp_path):
    p = tmp_path / "empty.fastq"
    p.write_bytes(b"")
    bases, quals, offsets, names = read_fastq(p)
    assert len(bases) == 0 and list(offsets) == [0] and names == []
    p = tmp_path / "lower.fastq"
    p.write_text("@a\nacgtNn\n+a\nIIIII!\n\n@b\nGG\n+\n#~\n")
    bases, quals, offsets, names = read_fastq(p)
    assert bases.tobytes() == b"ACGTNNGG" and list(quals) == [40] * 5 + [0, 2, 93] and names == ["a", "b"]
    for bad, what in (("a\nACGT\n+\nIIII\n", "'@'"), ("@a\nACGT\n+\nIII\n", "length"), ("@a\nACGT\nTTTT\n+\nIIIIIIII\n", "'+'"),
                      ("@a\nACGT\n+\n", "truncated"), ("@a\nAC\n+\nI \n", "Phred")):
        p = tmp_path / "bad.fastq"
        p.write_text(bad)
        with pytest.raises(sd.SdgpuError) as ei:
            read_fastq(p)
        assert what in str(ei.value) and "record 1" in str(ei.value)
    with pytest.raises(sd.SdgpuError):
        read_fastq(tmp_path / "missing.fastq")
