"""The `.myrtree` byte format (tax/store.rs:342-454): the library's reader / writer
(myrio_b200/csrc/myrtree.cu, through the C ABI) against the CPU oracle (oracle/myrtree_oracle.c) and
against a THIRD statement of the bincode layout assembled here with `struct` -- CPU only, no GPU needed
for the byte format.  The GPU half (K2's counts written back into a store, cached counts feeding the
index) is in test_gpu_myrtree.py.

Ported reference test: store.rs:465-506 `to_and_from_bytes_round_trip_test` (100 payloads "ACTG", 4 count
vectors {1: 2x, 10: 5, 200: 3} with rescale 0.1, k_precomputed [1, 2, 3, 4], one branch over 100 leaves).
"""
import os
import struct

import numpy as np
import pytest

import oracle
from myrio_b200 import api, synth

HERE = os.path.dirname(os.path.abspath(__file__))
MAGIC = bytes([77, 89, 82, 73, 79, 45, 206, 168])
IUPAC = {c: i for i, c in enumerate("-TGKCYSBAWRDMHVN")}


def codes(s):
    return np.array([IUPAC[c] for c in s], np.uint8)


# ---- third statement of the layout: bincode 2 fixint LE by hand ------------------------------------------------

def b_str(s):
    e = s.encode()
    return struct.pack("<Q", len(e)) + e


def b_seq(sym):  # ASSUMED layout: u64 symbols + Vec<usize> Lsb0 words
    c = codes(sym)
    nw = (len(c) + 15) // 16
    words = [0] * nw
    for i, v in enumerate(c):
        words[i // 16] |= int(v) << (4 * (i % 16))
    return struct.pack("<QQ", len(c), nw) + b"".join(struct.pack("<Q", w) for w in words)


def b_payload(sym, count_vecs):
    out = b_seq(sym) + struct.pack("<Q", len(count_vecs))
    for keys, vals, rescale in count_vecs:
        out += struct.pack("<Q", len(keys)) + b"".join(struct.pack("<Q", k) for k in keys)
        out += struct.pack("<Q", len(vals)) + b"".join(struct.pack("<H", v) for v in vals)
        out += struct.pack("<f", rescale)
    return out


def b_leaf(name, pid):
    return b"\x01" + b_str(name) + struct.pack("<Q", pid)


def b_branch(name, children_bytes):
    return b"\x00" + b_str(name) + struct.pack("<Q", len(children_bytes)) + b"".join(children_bytes)


# ---- the reference's own test --------------------------------------------------------------------------------

def reference_round_trip_store():
    n = 100
    kind = [0] + [1] * n
    arg = [n] + list(range(n))
    names = ["test-branch"] + ["test-leaf"] * n
    seq = np.tile(codes("ACTG"), n)
    off = np.arange(n + 1, dtype=np.uint64) * 4
    f = api.MyrTreeFile.new("test", api.RANKS["Species"], kind, arg, names, seq, off)
    ro = np.arange(n + 1, dtype=np.uint64) * 3
    keys = np.tile(np.array([1, 10, 200], np.uint64), n)
    vals = np.stack([2 * np.arange(n), np.full(n, 5), np.full(n, 3)], axis=1).astype(np.uint16).reshape(-1)
    resc = np.full(n, 0.1, np.float32)
    for k in (1, 2, 3, 4):
        f.append_counts_host(k, ro, keys, vals, resc)
    content = oracle.StoreContent("test", 1, np.array(kind, np.uint8), np.array(arg, np.uint64), names, seq, off,
                                  [1, 2, 3, 4], [oracle.Csr(ro, keys, vals, resc) for _ in range(4)])
    return f, content


def assert_same(f: api.MyrTreeFile, c: oracle.StoreContent):
    t = f.tree()
    assert t["gene"] == c.gene and t["root_rank"] == c.root_rank and t["k_precomputed"] == c.k_precomputed
    assert (t["kind"] == c.kind).all() and (t["arg"] == c.arg).all() and t["names"] == c.names
    words, woff, boff = f.seqs()
    assert (boff == c.seq_off).all()
    ew, ewo = synth.pack_iupac4(c.seq_codes, c.seq_off)
    assert (woff == ewo).all() and (words == ew).all()
    for i, e in enumerate(c.counts):
        ro, keys, vals, resc = f.counts_host(i)
        assert (ro == e.row_off).all() and (keys == e.keys).all() and (vals == e.vals).all()
        assert (resc.view(np.uint32) == np.asarray(e.aux, np.float32).view(np.uint32)).all()


def test_reference_round_trip_test_ported():
    # store.rs:465-506: to_bytes(3) -> from_bytes, then gene / root_rank / k_precomputed / payloads equal
    f, content = reference_round_trip_store()
    data = f.to_bytes(3)
    g = api.MyrTreeFile.from_bytes(data)
    assert_same(g, content)
    assert g.info()["n_payloads"] == 100 and g.info()["n_k"] == 4
    # the oracle reads the library's bytes and the library reads the oracle's
    o = oracle.store_decode(data)
    assert o.gene == "test" and o.root_rank == 1 and o.k_precomputed == [1, 2, 3, 4]
    assert (o.seq_codes == content.seq_codes).all() and (o.counts[3].vals == content.counts[3].vals).all()
    assert_same(api.MyrTreeFile.from_bytes(oracle.store_encode(content, 3, threads=8)), content)


@pytest.mark.parametrize("threads", [1, 3, 8, 200])
def test_writer_bytes_equal_the_oracle(threads):
    # same zstd library, same chunk rule (store.rs:363-367: len < n ? len : len / n + 1) => identical files
    f, content = reference_round_trip_store()
    mine = f.to_bytes(6, threads)
    theirs = oracle.store_encode(content, 6, threads)
    assert mine == theirs
    hdr = oracle.store_section(mine, -1)
    n_chunks = struct.unpack_from("<Q", hdr, 16)[0]
    chunk_size = 100 if 100 < threads else 100 // threads + 1
    assert n_chunks == -(-100 // chunk_size)
    assert mine[:8] == MAGIC and struct.unpack_from("<Q", mine, 8)[0] == len(hdr)


def test_bincode_layout_against_hand_assembled_bytes():
    # a small store with a two-level tree, ragged sequences (17 symbols: two words), an ambiguity code and a gap,
    # two cached k; sections decompressed with the oracle's zstd binding and compared with struct-packed bytes
    kind = [0, 0, 1, 1, 0, 1]
    arg = [2, 2, 0, 1, 1, 2]
    names = ["k:Plantae", "g:Abies", "s:Abies alba", "s:Abies böhmii", "g:Picea", "s:Picea abies"]
    seqs = ["ACGTN-RYACGTACGTA", "ACG", ""]
    cv = [[([5, 9], [32, 64], 0.03125), ([0], [1], 1.0)],
          [([7], [65535], 0.5), ([1, 2, 3], [6, 7, 8], 0.25)],
          [([], [], 2.0), ([2 ** 63], [9], 1.5)]]
    seq_codes = np.concatenate([codes(s) for s in seqs])
    off = np.array([0, 17, 20, 20], np.uint64)
    f = api.MyrTreeFile.new("ITS", api.RANKS["Kingdom"], kind, arg, names, seq_codes, off)
    for i, k in enumerate((6, 18)):
        ro = np.zeros(4, np.uint64)
        ro[1:] = np.cumsum([len(cv[p][i][0]) for p in range(3)])
        keys = np.array([x for p in range(3) for x in cv[p][i][0]], np.uint64)
        vals = np.array([x for p in range(3) for x in cv[p][i][1]], np.uint16)
        f.append_counts_host(k, ro, keys, vals, np.array([cv[p][i][2] for p in range(3)], np.float32))
    data = f.to_bytes(6, threads=2)  # chunk_size = 3 // 2 + 1 = 2 -> chunks of 2 and 1 payloads
    core = b_str("ITS") + b_branch("k:Plantae", [
        b_branch("g:Abies", [b_leaf("s:Abies alba", 0), b_leaf("s:Abies böhmii", 1)]),
        b_branch("g:Picea", [b_leaf("s:Picea abies", 2)])]) + struct.pack("<I", 7) + struct.pack("<Q", 0)
    assert oracle.store_section(data, 0) == core
    chunk0 = struct.pack("<Q", 2) + b_payload(seqs[0], cv[0]) + b_payload(seqs[1], cv[1])
    chunk1 = struct.pack("<Q", 1) + b_payload(seqs[2], cv[2])
    assert oracle.store_section(data, 1) == chunk0
    assert oracle.store_section(data, 2) == chunk1
    hdr = oracle.store_section(data, -1)
    vals = struct.unpack("<" + "Q" * (len(hdr) // 8), hdr)
    assert vals[1] == len(core) and vals[2] == 2 and vals[4] == len(chunk0) and vals[6] == len(chunk1)
    assert vals[7:] == (2, 6, 18)
    assert len(data) == 16 + len(hdr) + vals[0] + vals[3] + vals[5]
    # and back
    g = api.MyrTreeFile.from_bytes(data)
    t = g.tree()
    assert t["names"] == names and t["kind"].tolist() == kind and t["arg"].tolist() == arg and t["root_rank"] == 7
    assert g.counts_host(1)[1].tolist() == [0, 1, 2, 3, 2 ** 63]
    assert g.counts_host(0)[2].tolist() == [32, 64, 65535]
    assert g.seqs()[2].tolist() == [0, 17, 20, 20]


def test_golden_file_written_by_the_library():
    # tests/golden/tiny.myrtree was produced by the writer (tests/golden/make_golden.py) and is re-read here;
    # its sections still equal the hand-assembled bytes, whatever zstd version reads them
    path = os.path.join(HERE, "golden", "tiny.myrtree")
    data = open(path, "rb").read()
    g = api.MyrTreeFile.load_from_file(path)
    t = g.tree()
    assert t["gene"] == "trnH-psbA" and t["root_rank"] == api.RANKS["Family"] and t["k_precomputed"] == [6, 18]
    assert t["names"] == ["f:Pinaceae", "g:Abies", "s:Abies alba", "s:Abies alba", "g:Picea", "s:Picea abies"]
    assert t["kind"].tolist() == [0, 0, 1, 1, 0, 1] and t["arg"].tolist() == [2, 2, 0, 1, 1, 2]
    seqs = ["ACGTACGTTTGACCA", "ACGTACGTTTGACCANNRY-ACGT", "GGGG"]
    assert oracle.store_section(data, 0) == b_str("trnH-psbA") + b_branch("f:Pinaceae", [
        b_branch("g:Abies", [b_leaf("s:Abies alba", 0), b_leaf("s:Abies alba", 1)]),
        b_branch("g:Picea", [b_leaf("s:Picea abies", 2)])]) + struct.pack("<I", 3) + struct.pack("<Q", 0)
    o = oracle.store_decode(data)
    assert (o.seq_codes == np.concatenate([codes(s) for s in seqs])).all()
    # the cached counts are the oracle's counts of these sequences (k = 6 and 18, 32 resamples, seed 0)
    for i, k in enumerate((6, 18)):
        e = oracle.count_leaves(o.seq_codes, o.seq_off, k, 32)
        ro, keys, vals, resc = g.counts_host(i)
        assert (ro == e.row_off).all() and (keys == e.keys).all() and (vals == e.vals).all()
        assert (resc.view(np.uint32) == e.aux.view(np.uint32)).all()
    # re-encoding the decoded content reproduces the file's sections
    again = g.to_bytes(6, threads=1)
    for s in (0, 1):
        assert oracle.store_section(again, s) == oracle.store_section(data, s)


def test_malformed_input_is_an_error_not_a_crash():
    f, _ = reference_round_trip_store()
    data = f.to_bytes(1, threads=4)
    for bad in (b"", b"MYRIO", b"NOTMYRIO" + data[8:], data[:40], data[:-1], data[:len(data) // 2],
                data[:8] + struct.pack("<Q", 2 ** 60) + data[16:]):
        with pytest.raises(api.MyrioError) as e:
            api.MyrTreeFile.from_bytes(bad)
        assert e.value.code == 11, bad[:16]
        with pytest.raises(ValueError):
            oracle.store_decode(bad)
    # flipped bytes inside a compressed section: zstd or bincode reports it
    broken = bytearray(data)
    for i in range(len(data) - 60, len(data) - 40):
        broken[i] ^= 0x5A
    with pytest.raises(api.MyrioError):
        api.MyrTreeFile.from_bytes(bytes(broken))
    # a store without payloads cannot be serialised (par_chunks(0) panics in the reference)
    empty = api.MyrTreeFile.new("g", 1, [0], [0], ["root"], np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    with pytest.raises(api.MyrioError) as e:
        empty.to_bytes()
    assert e.value.code == 7
    # tree / payload consistency is checked when a store is assembled
    with pytest.raises(api.MyrioError):
        api.MyrTreeFile.new("g", 1, [0, 1], [1, 5], ["r", "l"], codes("ACGT"), np.array([0, 4], np.uint64))
    with pytest.raises(api.MyrioError):
        api.MyrTreeFile.new("g", 9, [1], [0], ["l"], codes("ACGT"), np.array([0, 4], np.uint64))


def test_shrink_and_file_round_trip(tmp_path):
    f, content = reference_round_trip_store()
    p = tmp_path / "x.myrtree"
    f.save_to_file(str(p), 3)
    g = api.MyrTreeFile.load_from_file(str(p))
    assert_same(g, content)
    g.shrink()  # store.rs:259-262
    assert g.k_precomputed == [] and g.info()["n_k"] == 0
    h = api.MyrTreeFile.from_bytes(g.to_bytes(3))
    assert h.info()["n_k"] == 0 and h.info()["n_payloads"] == 100
    with pytest.raises(api.MyrioError):
        api.MyrTreeFile.load_from_file(str(tmp_path / "missing.myrtree"))
