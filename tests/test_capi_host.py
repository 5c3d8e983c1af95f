"""CPU tier: the C-ABI library loads, exports every symbol include/vp8r.h declares, and the host
logic (parser errors, key-frame scan, checksum, transforms KATs) behaves.  No compute calls."""
import ctypes as C
import os
import re
import random

import pytest

import helpers


def test_header_and_library_agree(built):
    from vp8_b200 import _capi
    hdr = open(os.path.join(helpers.ROOT, "include", "vp8r.h")).read()
    declared = set(re.findall(r"VP8R_API\s+[\w \*]+?\b(vp8r_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(helpers.LIB_SO)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in vp8r.h but not exported"
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)


def test_struct_layout(built):
    from vp8_b200 import _capi
    assert C.sizeof(_capi.MbInfo) == 32
    assert C.sizeof(_capi.FrameHdr) == 8 + 16 + 48 + 32
    assert C.sizeof(_capi.TokenHdr) == 1152


def test_engine_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import vp8_b200
    with pytest.raises(vp8_b200._capi.Vp8rError) as e:
        vp8_b200.Engine(0)
    assert "no CPU fallback" in str(e.value)


def test_parser_errors(built):
    import vp8_b200
    from vp8_b200._capi import Vp8rError
    _, payloads = vp8_b200.read_ivf(helpers.vectors()[0])
    key = payloads[0]
    p = vp8_b200.Parser()
    with pytest.raises(Vp8rError) as e:  # inter frame before any key frame (VP8R_ERR_STATE)
        p.parse(bytes([key[0] | 1]) + key[1:])
    assert e.value.code == 5
    with pytest.raises(Vp8rError) as e:  # bad start code: ensure() in bitstream_parser.cc:28-29
        p.parse(key[:3] + b"\x00\x00\x00" + key[6:])
    assert e.value.code == 2
    with pytest.raises(Vp8rError) as e:  # experimental version: bitstream_parser.cc:24-25
        p.parse(bytes([key[0] | (4 << 1)]) + key[1:])
    assert e.value.code == 3
    with pytest.raises(Vp8rError) as e:  # truncated: the reference throws std::out_of_range
        p.parse(key[:len(key) // 3])
    assert e.value.code == 4
    with pytest.raises(Vp8rError):
        p.parse(b"")
    # and the parser still works afterwards
    d = p.parse(key).desc()
    assert d.hdr.key_frame == 1 and d.hdr.mb_cols == 11 and d.hdr.mb_rows == 9


def test_key_frame_scan(built):
    import vp8_b200
    lib = vp8_b200._capi.load()
    _, payloads = vp8_b200.read_ivf(os.path.join(helpers.VEC_DIR, "vp80-02-inter-1402.ivf"))
    flags = [lib.vp8r_is_key_frame(p, len(p)) for p in payloads]
    assert flags[0] == 1 and sum(flags) >= 1 and 0 in flags


def test_checksum_host(built):
    import vp8_b200
    lib = vp8_b200._capi.load()
    w, h = 7, 5
    n = helpers.i420_bytes(w, h)
    data = bytes((i * 37 + 11) & 255 for i in range(n))
    s1 = sum(data) & 0xFFFFFFFF
    s2 = sum((i + 1) * b for i, b in enumerate(data)) & 0xFFFFFFFF
    assert lib.vp8r_checksum_i420(data, w, h) == (s2 << 32) | s1


def test_transform_kats(built):
    """The reference's own unit tests (test/dct_test.h:16-71): a DC-only block inverse-transforms
    to (dc+4)>>3 everywhere; here checked on the oracle's IDCT (the CUDA IDCT is checked against
    the oracle on the GPU)."""
    orc = helpers.Oracle()
    rng = random.Random(7122)
    for _ in range(100):
        dc = rng.randrange(-2048, 2048)
        blk = (C.c_int16 * 16)(dc, *([0] * 15))
        orc.lib.oracle_idct4x4(blk)
        assert list(blk) == [(dc + 4) >> 3] * 16
    # IWHT of a DC-only block: every output (dc+3)>>3
    for _ in range(100):
        dc = rng.randrange(-2048, 2048)
        blk = (C.c_int16 * 16)(dc, *([0] * 15))
        orc.lib.oracle_iwht4x4(blk)
        assert list(blk) == [(dc + 3) >> 3] * 16
    orc.close()


def test_parsed_frame_accounting(built):
    """coef_mask / coef_offset / n_payload_blocks are mutually consistent for every frame of a
    stream that uses SPLIT motion vectors."""
    import vp8_b200
    ivf = helpers.synth_stream("--width 176 --height 144 --frames 6 --seed 11 --pct-split 40")
    _, payloads = vp8_b200.read_ivf(ivf)
    p = vp8_b200.Parser()
    for pl in payloads:
        fr = p.parse(pl)
        d = fr.desc()
        n_mb = d.hdr.mb_cols * d.hdr.mb_rows
        at = 0
        coef = split = inter = 0
        for i in range(n_mb):
            mb = d.mbs[i]
            is_split = (mb.flags & 1) and ((mb.flags >> 3) & 7) == 4
            if is_split:
                assert mb.aux[0] == at
                at += 2
                split += 1
            inter += mb.flags & 1
            assert mb.coef_offset == at
            k = bin(mb.coef_mask).count("1")
            at += k
            coef += k
            assert mb.coef_mask < (1 << 25)
            if not (mb.flags & 0x100):
                assert not (mb.coef_mask & 1)  # no Y2 block without has_y2
        assert at == (d.hdr.intra_levels_at if d.hdr.n_intra_levels else d.hdr.n_payload_blocks) and coef == d.hdr.n_coef_blocks
        assert split == d.hdr.n_split_mbs and inter == d.hdr.n_inter_mbs
        fr.close()


def test_parser_survives_corrupted_streams(built):
    """Bit flips, truncations and garbage: the parser must return a status (never crash or hang), and
    whatever it accepts must still satisfy the array invariants the kernels rely on."""
    import vp8_b200
    from vp8_b200._capi import Vp8rError
    rng = random.Random(99)
    base = helpers.synth_stream("--width 96 --height 80 --frames 6 --seed 5 --log2-parts 2 --pct-split 30 --pct-intra 20")
    _, payloads = vp8_b200.read_ivf(base)
    accepted = rejected = 0
    for trial in range(300):
        p = vp8_b200.Parser()
        for k, pl in enumerate(payloads):
            data = bytearray(pl)
            mode = rng.randrange(4)
            if mode == 0:
                for _ in range(rng.randint(1, 8)):
                    data[rng.randrange(len(data))] ^= 1 << rng.randrange(8)
            elif mode == 1:
                data = data[:rng.randrange(1, len(data))]
            elif mode == 2:
                start = rng.randrange(len(data))
                for i in range(start, min(len(data), start + 32)):
                    data[i] = rng.randrange(256)
            try:
                fr = p.parse(bytes(data))
            except Vp8rError as e:
                assert 1 <= e.code <= 7
                rejected += 1
                continue
            accepted += 1
            d = fr.desc()
            n_mb = d.hdr.mb_cols * d.hdr.mb_rows
            assert 0 < n_mb <= 1 << 20
            limit = d.hdr.intra_levels_at if d.hdr.n_intra_levels else d.hdr.n_payload_blocks
            for i in range(n_mb):
                mb = d.mbs[i]
                assert mb.coef_offset + bin(mb.coef_mask).count("1") <= limit
                assert ((mb.flags >> 11) & 63) <= 63 and mb.coef_mask < (1 << 25)
                if (mb.flags & 1) and ((mb.flags >> 3) & 7) == 4:
                    assert mb.aux[0] + 2 <= limit
                if not (mb.flags & 1) and ((mb.flags >> 3) & 7) == 4:
                    assert all(((mb.aux[b >> 3] >> ((b & 7) * 4)) & 15) <= 9 for b in range(16))
            fr.close()
    assert accepted > 100 and rejected > 20


def test_deferred_token_parse_matches_full_parse(built):
    """vp8r_parser_set_defer_tokens: the first-partition half of the parse is unchanged (modes, motion
    vectors, loop-filter levels, intra levels), coefficient fields are left for the device, and the DCT
    partitions + the frame's token probabilities are attached to the payload."""
    import ctypes as C
    import vp8_b200
    from vp8_b200._capi import TokenHdr
    INNER, SKIP = 0x20000, 0x40000
    streams = [open(v, "rb").read() for v in helpers.vectors()[:6]]
    streams.append(helpers.synth_stream("--width 176 --height 144 --frames 6 --seed 3 --log2-parts 3 --pct-split 30 --pct-intra 20"))
    for data in streams:
        _, payloads = vp8_b200.read_ivf(data)
        full, lazy = vp8_b200.Parser(), vp8_b200.Parser()
        lazy.set_defer_tokens(True)
        for pl in payloads:
            a, b = full.parse(pl), lazy.parse(pl)
            da, db = a.desc(), b.desc()
            assert db.hdr.tokens_deferred == 1 and da.hdr.tokens_deferred == 0
            for f in ("width", "height", "mb_cols", "mb_rows", "key_frame", "version", "show_frame", "filter_type",
                      "loop_filter_level", "sharpness_level", "refresh_last", "refresh_golden", "refresh_altref",
                      "copy_to_golden", "copy_to_altref", "n_inter_mbs", "n_split_mbs", "n_intra_levels"):
                assert getattr(da.hdr, f) == getattr(db.hdr, f), f
            assert bytes(da.hdr.dq) == bytes(db.hdr.dq)
            n_mb = da.hdr.mb_cols * da.hdr.mb_rows
            for i in range(n_mb):
                ma, mb = da.mbs[i], db.mbs[i]
                assert (ma.flags & ~INNER) == (mb.flags & ~(INNER | SKIP))
                assert mb.coef_mask == 0 and mb.coef_offset == 0
                if mb.flags & SKIP:
                    assert ma.coef_mask == 0
                bpred_or_split = ((ma.flags >> 3) & 7) == 4
                assert bool(mb.flags & INNER) == bpred_or_split
                assert bool(ma.flags & INNER) == (bpred_or_split or ma.coef_mask != 0)
                assert tuple(ma.mv) == tuple(mb.mv)
                if (ma.flags & 1) and bpred_or_split:  # SPLIT: same 16 motion vectors
                    assert da.payload[ma.aux[0] * 16: ma.aux[0] * 16 + 32] == db.payload[mb.aux[0] * 16: mb.aux[0] * 16 + 32]
                else:
                    assert tuple(ma.aux) == tuple(mb.aux)
            th = C.cast(C.addressof(db.payload.contents) + db.hdr.tokens_at * 32, C.POINTER(TokenHdr)).contents
            assert th.n_parts in (1, 2, 4, 8)
            tag = pl[0] | pl[1] << 8 | pl[2] << 16
            first = (10 if not (tag & 1) else 3) + (tag >> 5)
            total = len(pl) - first - 3 * (th.n_parts - 1)
            got = sum(th.part_size[k] for k in range(th.n_parts))
            # a last partition shorter than 2 bytes is treated as absent, like the host reader does
            assert got == total or (th.part_size[th.n_parts - 1] == 0 and total - got < 2)
            raw = C.string_at(C.addressof(th) + C.sizeof(TokenHdr), th.raw_bytes)
            at = first + 3 * (th.n_parts - 1)
            for k in range(th.n_parts):
                assert raw[th.part_off[k]: th.part_off[k] + th.part_size[k]] == pl[at: at + th.part_size[k]]
                at += th.part_size[k]
            a.close()
            b.close()


def test_deferred_parse_falls_back_for_huge_frames(built):
    """Frames with more than 65536 macroblocks are parsed on the host even when deferral was asked for
    (the device parse kernel keeps two bytes per macroblock in shared memory)."""
    import vp8_b200
    data = helpers.synth_stream("--width 4112 --height 4112 --frames 1 --seed 2 --log2-parts 1 --pct-skip 90")
    _, payloads = vp8_b200.read_ivf(data)
    full, lazy = vp8_b200.Parser(), vp8_b200.Parser()
    lazy.set_defer_modes(True)
    a, b = full.parse(payloads[0]), lazy.parse(payloads[0])
    da, db = a.desc(), b.desc()
    assert da.hdr.mb_cols * da.hdr.mb_rows == 257 * 257 > 65536
    assert db.hdr.tokens_deferred == 0 and db.hdr.modes_deferred == 0
    assert db.hdr.n_payload_blocks == da.hdr.n_payload_blocks and db.hdr.n_coef_blocks == da.hdr.n_coef_blocks
    n = da.hdr.mb_cols * da.hdr.mb_rows
    assert all(da.mbs[i].flags == db.mbs[i].flags and da.mbs[i].coef_mask == db.mbs[i].coef_mask for i in range(0, n, 97))
    a.close()
    b.close()


def test_batch_payload_pointers_in_place(built):
    """BatchDecoder hands `bytes` payloads to the parser without copying; other buffer types are copied."""
    import ctypes as C
    from vp8_b200.batch import BatchDecoder
    p0, p1 = b"\x01\x02\x03\x04", bytearray(b"\x09\x08\x07")
    ptrs, keep = BatchDecoder._payload_pointers([p0, p1, b""], [0, 1, 2])
    assert len(ptrs) == 3 and len(keep) == 1
    assert C.string_at(ptrs[0], 4) == p0 and C.string_at(ptrs[1], 3) == bytes(p1)
    assert ptrs[0] == C.cast(C.c_char_p(p0), C.c_void_p).value  # the bytes object's own buffer
