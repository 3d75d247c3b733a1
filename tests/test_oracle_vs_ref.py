"""The C restatement against the UNMODIFIED reference (oracle/_ref) on fresh seeded inputs, every stage bit-exact.
Skipped where oracle/_ref is not built (it needs /root/reference at build time)."""
import numpy as np
import pytest

from oracle import port, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libmercury_ref.so not built")

THRESH_DB = [-10, -7.5, -6, -4.5, -3.5, -2.5, -1.5, -0.5, 0.5, 1.5, 3, 4, 6.5, 7.5, 9, 12.5, 13.5]


@pytest.mark.parametrize("cfg", list(range(17)))
def test_every_stage_bit_exact(cfg):
    iters = 20 if cfg == 16 else 50
    r, p = ref.Ref(cfg, iters), port.Port(cfg, iters)
    gr, gp = dict(r.geom), dict(p.geom)
    for k in ("buffer_Nsymb", "total_frame_size", "Cwidth", "Vwidth", "dwidth"):
        gr.pop(k), gp.pop(k)
    assert gr == gp
    tr, tp = r.tables(), p.tables()
    for k in ("carrier_type", "pilot_seq", "scrambler", "constellation"):
        assert np.array_equal(tr[k], tp[k]), k
    lr, lp = ref.ldpc_tables(ref.RATE_OF_CONFIG[cfg]), p.ldpc_tables()
    for k in ("C", "V", "d", "Enc"):
        assert np.array_equal(lr[k], lp[k]), k
    rng = np.random.default_rng(4242 + cfg)
    # +3 dB (decodes), +0.5 dB (many iterations), -3 dB (fails: exercises the I+1 path and garbage payloads)
    # ZF modes only decode as a hard-decision pass-through far above the nominal threshold: +16 dB exercises
    # the re-encode SNR report (telecom_system.cc:1376-1400)
    for off in ((3.0, 0.5, -3.0) if cfg < 15 else (16.0, 11.5, -3.0)):
        pl = rng.integers(0, 256, r.frame_bytes)
        xr, ar = r.tx_baseband(pl, True)
        xp, ap = p.tx_baseband(pl, True)
        assert np.array_equal(xr, xp)
        for k in ar:
            assert np.array_equal(ar[k], ap[k]), k
        sigma = 10 ** (-(THRESH_DB[cfg] + off) / 20) * 16
        x = xr + (rng.standard_normal(xr.size) + 1j * rng.standard_normal(xr.size)) * sigma / np.sqrt(2)
        a, b = r.rx_tail(x), p.rx_tail(x)
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (k, off)


def test_prng_and_crc():
    for seed in (0, 1, 5, 12345):
        assert np.array_equal(ref.ref_random(seed, 2000), port.port_random(seed, 2000))
    rng = np.random.default_rng(1)
    for n in (1, 2, 10, 75, 175):
        d = rng.integers(0, 256, n)
        assert ref.ref_crc16(d) == port.port_crc16(d)
