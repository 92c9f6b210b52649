"""CPU oracle for the phoneme_contrast training hot path.

TEST INFRASTRUCTURE ONLY. Nothing in ``phoneme_contrast_b200/`` may import this
package. The only legal callers are ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and there only as
the checker or as the reported CPU baseline -- never as the thing shipped.

Each function restates one piece of the reference's algorithm on the CPU and
cites the reference (or torchaudio) file:line it follows:

* ``mfcc_oracle``    -- torchaudio.transforms.MFCC as configured by
                        src/datasets/features.py:25-55 (numpy fp64 restatement
                        + a torch-op variant used for CPU timing)
* ``augment_oracle`` -- src/datasets/transforms.py:25-97,129-144 and
                        src/datasets/dataset.py:147-172 (mask / noise / gain
                        decisions and indices)
* ``supcon_oracle``  -- src/training/losses.py:26-86 (+ analytic gradient)
* ``nets_oracle``    -- src/models/phoneme_cnn.py:10-304 as torch functional ops
* ``optim_oracle``   -- clip_grad_norm_ (src/training/trainer.py:147-150) +
                        Adam with L2 weight decay (scripts/train.py:129-133)

Parity pinning: every module is checked against golden vectors produced by the
real reference (imported from /root/reference with the installed torchaudio
2.11.0) by ``tests/golden/make_golden.py``; see tests/test_oracle_golden.py.
The reference's own tests hold no numeric known-answer vectors for this path
(SURVEY.md section 8c), so the goldens generated from the running reference are the
pin. The reference pins torchaudio 2.7.0 (uv.lock); the container has 2.11.0,
whose MFCC code path is the one the goldens were generated with.
"""
