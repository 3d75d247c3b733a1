mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python -c "
import numpy as np, mercury_b200 as mb
from oracle import ref
from tests import frontend_cases as fc
ts = mb.TelecomSystemB200(0)
for cfg in (8, 16, 0):
    r = ref.Ref(cfg, 50)
    ts.load_configuration(cfg, 50)
    n = r.capture_samples()
    rng = np.random.default_rng(cfg)
    caps = []
    for df, sigma in ((12.1, 0.1), (-11.4, 0.1), (0.0, 0.6), (3.0, 0.02)):
        tx = r.transmit_byte(rng.integers(0, 256, r.frame_bytes)); d = int(rng.integers(6000, 30000))
        cap = np.zeros(n); cap[d:d + tx.size] += tx
        caps.append((fc.freq_shift(cap, df) + rng.normal(0, sigma, n)).astype(np.float32).astype(np.float64))
    caps = np.stack(caps)
    for en in (True, False):
        ts.set_coarse_freq_sync(en)
        st = mb.new_receive_stats(len(caps)); st['delay_of_last_decoded_message'][:] = -1
        p, s, _ = ts.receive_byte_batch(caps, st)
        print(cfg, en, s['sync_trials'].tolist(), s['message_decoded'].tolist())
" > gpurun_out/r4_sanitize_cfs.log 2>&1
tail -12 gpurun_out/r4_sanitize_cfs.log
