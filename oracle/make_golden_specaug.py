"""TEST INFRASTRUCTURE ONLY — golden vectors of SpecAugment: the UNMODIFIED reference module
(lcasr/utils/augmentation.py:10-104) on CPU under a fixed torch seed.  The uniform draws it consumes are recorded
(torch.rand is wrapped while the module runs) so that the oracle restatement and the CUDA kernel can be checked on the
same masks independently of the random-number generator.
    python oracle/make_golden_specaug.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import lcasr_oracle as O  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
# name -> (B, F, T, lengths or None, module kwargs)
CASES = {
    "specaug_iid_minp": (3, 80, 1000, [1000, 730, 512], dict(n_time_masks=6, n_freq_masks=2, freq_mask_param=27, min_p=0.05, max_p=1.0)),
    "specaug_iid_zero": (2, 80, 513, None, dict(n_time_masks=2, n_freq_masks=3, freq_mask_param=15, time_mask_param=40, max_p=0.5,
                                                 zero_masking=True)),
    "specaug_shared": (4, 80, 256, None, dict(n_time_masks=3, n_freq_masks=1, freq_mask_param=10, time_mask_param=30, iid_masks=False)),
    "specaug_no_time": (2, 80, 64, [64, 40], dict(n_time_masks=0, n_freq_masks=2, freq_mask_param=20)),
}


def main():
    load_reference()
    from lcasr.utils.augmentation import SpecAugment
    for k, (name, (B, F, T, lengths, kw)) in enumerate(CASES.items()):
        spec = O.synth_input(B, T, F, seed=77 + k) + 0.3
        lens = None if lengths is None else torch.tensor(lengths)
        draws, real_rand = [], torch.rand

        def rec(*a, **kwargs):
            r = real_rand(*a, **kwargs)
            draws.append(r.detach().clone().reshape(-1))
            return r

        torch.manual_seed(1000 + k)
        torch.rand = rec
        try:
            ref = SpecAugment(**kw)(spec.clone(), lens)
        finally:
            torch.rand = real_rand
        tp, fp = O.specaug_params(T, F, kw["n_time_masks"], kw["n_freq_masks"], kw["freq_mask_param"], kw.get("time_mask_param", -1),
                                  kw.get("min_p", -1), kw.get("max_p", 1.0))
        nt = kw["n_time_masks"] if tp >= 1 else 0
        nf = kw["n_freq_masks"] if fp >= 1 else 0
        assert len(draws) == 2 * (nt + nf), (len(draws), nt, nf)
        per = draws[0].numel()
        u = torch.stack(draws).reshape(nt + nf, 2, per).numpy().astype(np.float32)
        u_time, u_freq = u[:nt], u[nt:]
        mine = O.spec_augment(spec.numpy(), lengths, tp, u_time, fp, u_freq, zero_masking=kw.get("zero_masking", False))
        changed = float((ref.numpy() != spec.numpy()).mean())
        err = float(np.abs(mine - ref.numpy()).max())
        print(f"{name}: [{B},{F},{T}] time {nt}x{tp} freq {nf}x{fp} draws/mask {per}: {100 * changed:.1f}% masked; "
              f"oracle-vs-reference max-abs {err:.2e}")
        assert err < 1e-6 and np.array_equal(mine != spec.numpy(), ref.numpy() != spec.numpy())
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), B=B, F=F, T=T, lengths=np.array(lengths if lengths else [], dtype=np.int64),
                            kwargs=repr(kw), input_seed=77 + k, torch_seed=1000 + k, time_param=tp, freq_param=fp, u_time=u_time, u_freq=u_freq,
                            out=ref.numpy().astype(np.float32))


if __name__ == "__main__":
    main()
