"""Randomised stress of the order / index kernels against their torch definitions, bit for bit:
  * cng_merge_sort vs torch.sort(stable) of cat([fine, coarse]) -- random S, value ranges over many binades, heavy ties, sorted and
    unsorted coarse lists;
  * cng_sample_pdf / cng_resample_from_coarse vs oracle.resample_pdf -- random bin counts, peaked and flat weights, u at the ends.
    python tools/gpu/stress_c5.py [rounds]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from conditioned_nerf_gan_b200 import ops
from oracle import nerf_path as oracle

dev = "cuda"


def run(rounds: int, seed: int = 1234) -> int:
  """number of mismatching rounds"""
  g = torch.Generator().manual_seed(seed)
  bad = 0
  for it in range(rounds):
      S = int(torch.randint(2, 257, (1,), generator=g))
      n = int(torch.randint(1, 400, (1,), generator=g))
      mode = it % 5
      if mode == 0:
          t_c, t_f = torch.rand((n, S), generator=g) * 0.24 + 0.88, torch.rand((n, S), generator=g) * 0.24 + 0.88
      elif mode == 1:
          t_c, t_f = torch.exp(torch.rand((n, S), generator=g) * 30 - 15), torch.exp(torch.rand((n, S), generator=g) * 30 - 15)
      elif mode == 2:                                           # heavy ties: values on a coarse grid
          t_c, t_f = torch.randint(0, 9, (n, S), generator=g).float() / 8, torch.randint(0, 9, (n, S), generator=g).float() / 8
      elif mode == 3:
          t_c, t_f = torch.randn((n, S), generator=g) * 5, torch.randn((n, S), generator=g) * 5
      else:                                                     # fine samples concentrated in one coarse interval
          t_c = torch.rand((n, S), generator=g) * 1.7 + 0.25
          t_f = t_c[:, S // 2: S // 2 + 1] + torch.rand((n, S), generator=g) * 1e-4
      t_c = torch.sort(t_c, dim=1).values
      if it % 7 == 3 and n > 1:
          t_c[0] = t_c[0].flip(0)                               # one unsorted coarse ray: the generic path
      order, t_sorted = ops.merge_sort(t_f.to(dev).unsqueeze(0), t_c.to(dev).unsqueeze(0), want_sorted=True)
      ref_t, ref_i = torch.sort(torch.cat([t_f, t_c], dim=1), dim=1, stable=True)
      ok = torch.equal(t_sorted.cpu().reshape(n, 2 * S), ref_t)
      if mode != 3:                                             # signed zeros compare equal for torch: order checked by value there
          ok = ok and torch.equal(order.cpu().long().reshape(n, 2 * S), ref_i)
      if not ok:
          bad += 1
          print(f"MERGE MISMATCH round {it}: S={S} n={n} mode={mode}")
      # sample_pdf
      M = int(torch.randint(1, 300, (1,), generator=g))
      K = int(torch.randint(1, 300, (1,), generator=g))
      bins = torch.sort(torch.rand((n, M + 1), generator=g) * 1.7 + 0.25, dim=1).values
      w = torch.rand((n, M), generator=g)
      if it % 3 == 0:
          w = w ** 12                                           # peaked
      if it % 4 == 1:
          w[:, ::2] = 0                                         # empty bins: denominators under eps
      u = torch.rand((n, K), generator=g)
      u[:, 0] = 0.0
      if K > 1:
          u[:, 1] = 1.0 - 2.0 ** -24
      s_ref, i_ref, _, _ = oracle.resample_pdf(bins, w, u)
      s, i = ops.sample_pdf(bins.to(dev), w.to(dev), u.to(dev), want_inds=True)
      if not (torch.equal(i.cpu(), i_ref) and torch.equal(s.cpu(), s_ref)):
          bad += 1
          d = (s.cpu() - s_ref).abs().max().item()
          print(f"SAMPLE_PDF MISMATCH round {it}: n={n} M={M} K={K} inds equal {torch.equal(i.cpu(), i_ref)} max |ds| {d:.3e}")
  return bad


if __name__ == "__main__":
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    bad = run(rounds)
    print(f"stress_c5: {rounds} rounds, {bad} mismatches")
    sys.exit(1 if bad else 0)
