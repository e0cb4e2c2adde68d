"""CPU oracle for the Imagine360 dual-branch denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``imagine360_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may.

It is a plain-PyTorch, functional (state-dict in, tensor out) restatement of the reference's
algorithm for the path named in BASELINE.json -- each function cites the reference file:line it
follows.  The reference is pure Python over torch/xformers/kornia; the arithmetic that lives in the
absent third-party packages is restated from their published semantics:

* xformers 0.0.28.post1 ``memory_efficient_attention`` (requirements.txt:4):
  softmax(q k^T / sqrt(d) + bias) v,
* kornia (unpinned, requirements.txt:21) ``remap`` / ``gaussian_blur2d`` / ``create_meshgrid``.

Pinning: the reference holds NO golden vectors or tests for this path, so the oracle
is pinned against outputs of the unmodified reference modules imported in the build container
(tools/ref_shim.py + tools/make_golden.py -> tests/golden/*.pt, checked by tests/test_oracle_golden.py).
At the two third-party boundaries (kornia, xformers) parity is "unpinned" in the sense of §8(c):
no reference test fixes their results; the restatement follows the public library semantics.
"""
