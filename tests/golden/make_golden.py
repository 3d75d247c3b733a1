#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref) -- run in the build container only.

The reference ships no golden vectors, so the pins are outputs of the reference itself:

  rx_mode<NN>.npz  one frame per CONFIG_0..16: random payload -> reference TX chain (transmit_byte bit chain +
                   baseband_test_EsN0 modulation chain) -> AWGN at the mode's FER<0.1 threshold + 2 dB
                   (include/common/common_defines.h:130-147) -> rounded to complex64 (the GPU path's input
                   type; the reference then runs on exactly those values widened to double) -> every stage of
                   the reference RX tail (telecom_system.cc:1132-1341) recorded.
  loopback_mode00.npz  BASELINE config #1: TX_TEST payload (telecom_system.cc:2050-2055) -> reference
                   transmit_byte(SINGLE_MESSAGE) pass-band -> placed in a capture buffer -> reference
                   receive_byte(); records the post-synchronisation baseband_data its hot path consumed
                   plus the bytes / stats it returned.
  kat.npz          PRNG / CRC / table known answers.

Usage: python tests/golden/make_golden.py   (needs oracle/_ref/libmercury_ref.so, i.e. /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.normpath(os.path.join(HERE, "..", "..")))
from oracle import ref  # noqa: E402

# Es/N0 (dB) for FER<0.1 per mode: include/common/common_defines.h:130-147
THRESH_DB = [-10, -7.5, -6, -4.5, -3.5, -2.5, -1.5, -0.5, 0.5, 1.5, 3, 4, 6.5, 7.5, 9, 12.5, 13.5]
ITERS = {16: 20}


def awgn(rng, n, esn0_db):
    """baseband_test_EsN0 normalisation (telecom_system.cc:141-153): sigma*sqrt(Nfft) per complex sample."""
    sigma = 10.0 ** (-esn0_db / 20.0) * 16.0
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (sigma / np.sqrt(2.0))


def main():
    for cfg in range(17):
        r = ref.Ref(cfg, ITERS.get(cfg, 50))
        rng = np.random.default_rng(1000 + cfg)
        payload = rng.integers(0, 256, r.frame_bytes).astype(np.int32)
        x, aux = r.tx_baseband(payload, want_aux=True)
        esn0 = THRESH_DB[cfg] + (2.0 if cfg < 15 else 14.0)  # ZF modes 15/16 only decode as pass-through
        xin = (x + awgn(rng, x.size, esn0)).astype(np.complex64)
        o = r.rx_tail(xin.astype(np.complex128))
        np.savez_compressed(
            os.path.join(HERE, f"rx_mode{cfg:02d}.npz"),
            config=cfg, ldpc_iters=r.ldpc_iters, esn0_db=esn0, payload=payload, x=xin,
            codeword=aux["codeword"].astype(np.uint8), info_bits=aux["info_bits"].astype(np.uint8),
            Y=o["Y"].astype(np.complex64), H=o["H"].astype(np.complex64), Z=o["Z"].astype(np.complex64),
            llr_demod=o["llr_demod"], llr_cw=o["llr_cw"], bits=o["bits"].astype(np.uint8),
            bytes=o["bytes"].astype(np.uint8), rx_payload=o["payload"].astype(np.uint8),
            iterations=o["iterations"], crc=o["crc"], all_zeros=o["all_zeros"], decoded=o["decoded"],
            snr=o["snr"], variance=o["variance"], mean_H=o["mean_H"],
        )
        print(f"mode {cfg:2d}: EsN0 {esn0:5.1f} dB iters {o['iterations']:2d} decoded {o['decoded']} "
              f"payload_ok {np.array_equal(o['payload'], payload)}")
        r.close()

    # BASELINE config #1: mode 0 pass-band loop-back through the reference's own receive_byte()
    r = ref.Ref(0, 50)
    payload = np.zeros(r.frame_bytes, np.int32)
    payload[0] = 1  # TX_TEST frame, counter = 0 (telecom_system.cc:2050-2055)
    pb = r.transmit_byte(payload)
    sym = r.Nofdm * r.interp_rate
    buf = np.zeros(r.Nofdm * r.buffer_Nsymb * r.interp_rate)
    delay = (r.preamble_nSymb + 2) * sym + 200
    rng = np.random.default_rng(77)
    buf += rng.standard_normal(buf.size) * 1e-4
    buf[delay:delay + pb.size] += pb
    o = r.receive_byte(buf)
    pre = r.preamble_nSymb * r.Nofdm
    bb = o["baseband"][pre:].astype(np.complex64)
    t = r.rx_tail(bb.astype(np.complex128))  # the tail replayed on the float-rounded capture
    np.savez_compressed(
        os.path.join(HERE, "loopback_mode00.npz"),
        payload=payload, x=bb, rx_payload_receive_byte=o["payload"].astype(np.uint8), decoded_receive_byte=o["decoded"],
        iterations_receive_byte=o["iterations"], snr_receive_byte=o["snr"], delay=o["delay"],
        rx_payload=t["payload"].astype(np.uint8), decoded=t["decoded"], iterations=t["iterations"], crc=t["crc"],
        snr=t["snr"], llr_cw=t["llr_cw"],
    )
    print("loopback mode 0: receive_byte decoded", o["decoded"], "iters", o["iterations"], "delay", o["delay"],
          "payload_ok", np.array_equal(o["payload"], payload), "| tail replay decoded", t["decoded"], "iters", t["iterations"])

    tabs = {}
    for cfg in (0, 8, 10, 13, 16):
        rr = ref.Ref(cfg, 50)
        t = rr.tables()
        tabs[f"pilot_mask_{cfg}"] = (t["carrier_type"] == 1).astype(np.uint8)
        tabs[f"pilot_seq_{cfg}"] = t["pilot_seq"]
        tabs[f"constellation_{cfg}"] = t["constellation"]
        tabs[f"scrambler_{cfg}"] = t["scrambler"].astype(np.uint8)
        rr.close()
    np.savez_compressed(
        os.path.join(HERE, "kat.npz"),
        random_seed5=ref.ref_random(5, 64), random_seed0=ref.ref_random(0, 64), random_seed1=ref.ref_random(1, 64),
        crc_1234=ref.ref_crc16(b"1234"), **tabs,
    )


if __name__ == "__main__":
    main()
