"""Minimal BAM decoder (SAM specification section 4) for the tests: BGZF is a series of gzip members, which Python's gzip
module reads transparently.  Test infrastructure."""
import gzip
import struct

_SEQ = "=ACMGRSVTWYHKDBN"
_CIG = "MIDNSHP=X"
_SZ = {"c": ("<b", 1), "C": ("<B", 1), "s": ("<h", 2), "S": ("<H", 2), "i": ("<i", 4), "I": ("<I", 4), "f": ("<f", 4), "d": ("<d", 8)}


def read_bam(path):
    """-> (header_text, [(name, length)], [record dict])"""
    with gzip.open(path, "rb") as f:
        d = f.read()
    assert d[:4] == b"BAM\x01", "not a BAM file"
    l_text = struct.unpack_from("<i", d, 4)[0]
    text = d[8:8 + l_text].rstrip(b"\0").decode()
    at = 8 + l_text
    n_ref = struct.unpack_from("<i", d, at)[0]
    at += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", d, at)[0]
        name = d[at + 4:at + 4 + l_name - 1].decode()
        length = struct.unpack_from("<i", d, at + 4 + l_name)[0]
        refs.append((name, length))
        at += 8 + l_name
    recs = []
    while at < len(d):
        block = struct.unpack_from("<i", d, at)[0]
        r = d[at + 4:at + 4 + block]
        at += 4 + block
        tid, pos, l_name, mapq, bin_, n_cig, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHiiii", r, 0)
        o = 32
        qname = r[o:o + l_name - 1].decode()
        o += l_name
        cig = "".join(f"{c >> 4}{_CIG[c & 15]}" for c in struct.unpack_from(f"<{n_cig}I", r, o)) or "*"
        o += 4 * n_cig
        seq = "".join(_SEQ[(r[o + i // 2] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        o += (l_seq + 1) // 2
        q = r[o:o + l_seq]
        qual = "*" if (l_seq == 0 or all(x == 0xFF for x in q)) else "".join(chr(x + 33) for x in q)
        o += l_seq
        tags = {}
        while o < len(r):
            tag, typ = r[o:o + 2].decode(), chr(r[o + 2])
            o += 3
            if typ == "A":
                val, o = chr(r[o]), o + 1
            elif typ in _SZ:
                fmt, sz = _SZ[typ]
                val, o = struct.unpack_from(fmt, r, o)[0], o + sz
            elif typ in "ZH":
                e = r.index(b"\0", o)
                val, o = r[o:e].decode(), e + 1
            elif typ == "B":
                sub, n = chr(r[o]), struct.unpack_from("<I", r, o + 1)[0]
                fmt, sz = _SZ[sub]
                val = (sub, list(struct.unpack_from(f"<{n}{fmt[1]}", r, o + 5)))
                o += 5 + n * sz
            else:
                raise ValueError(f"unknown tag type {typ}")
            tags[tag] = (typ if typ not in "cCsSiI" else "i", val)
        recs.append(dict(qname=qname, flag=flag, tid=tid, pos=pos, mapq=mapq, bin=bin_, cigar=cig, mtid=mtid, mpos=mpos, tlen=tlen,
                         seq=seq, qual=qual, tags=tags))
    return text, refs, recs
