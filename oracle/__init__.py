"""CPU oracle for the conditioned-NeRF-GAN rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``conditioned_nerf_gan_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the timed CPU arm.

The reference (zzhuolun/conditioned-nerf-gan) is pure Python on top of PyTorch, so the oracle
is a torch-CPU fp32 restatement of the reference algorithm, stage by stage, each function
citing the reference file:line it follows (paths relative to the reference checkout).

Parity pinning: the reference ships no tests / golden vectors, so the
oracle is pinned against outputs of the *unmodified reference itself* executed in the build
container: ``tests/golden/make_golden.py`` imports ``/root/reference/generators`` (with a
``matplotlib`` stub), records every stage tap on seeded inputs and replayed RNG draws, and the
committed ``tests/golden/*.npz`` fixtures are checked against this oracle by
``tests/test_oracle_golden.py``.
"""

from .nerf_path import *  # noqa: F401,F403
