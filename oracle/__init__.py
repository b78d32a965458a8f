"""CPU oracle for the MorpheuS render-and-loss hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy / torch-CPU) of the
reference algorithms on the hot path, each function citing the reference file:line it
follows.  It exists so that the CUDA product path in ``morpheus_b200`` can be checked
against something that is *not* itself.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product
package never does (``tests/test_abi.py::test_product_never_imports_oracle`` greps for that).

Pinning status (see DESIGN.md "Oracle"):
  * grid encoder  -- pinned against the *reference CUDA kernel itself*: ``oracle/build_ref.sh``
    compiles the unmodified ``external/encoders/gridencoder/src`` into ``oracle/_ref`` (travels to the GPU box), and
    ``tests/test_gpu_parity.py::test_grid_encode_bit_exact_vs_reference_kernel`` compares the product kernel with it bit for bit;
    ``oracle/grid.py`` is within 1e-6 of the product kernel, hence pinned transitively.  The reference's own ``grid.py`` host code
    is exercised unchanged by ``tests/test_reference_grid_binding.py``.
  * scene model (MLPs, encodings, codes, Laplace density, shading, normals, pose) -- pinned
    against the reference Python (``/root/reference/models/*.py`` imported unmodified on CPU)
    through ``tests/golden/scene_*.npz`` (script: ``tests/golden/make_scene_golden.py``).
  * nerfacc sampling / compositing -- **parity unpinned**: nerfacc is an un-vendored pip
    dependency with no version pin (docs/INSTALL.md:22-23) and is not installable offline.
    The compositing math is closed form and restated from the published nerfacc 0.5.x API
    semantics; the sampler is validated by invariants only.
  * SDS -- the scalar math (schedule, add-noise, CFG, w(t), loss) is closed form and pinned by ``tests/golden/sds_chain.npz`` (the reference
    method executed from its source); the networks are pinned against the reference ``UNetModel`` / ``Encoder`` classes with SEEDED RANDOM
    weights (``tests/golden/sds_nets.npz``): the Zero-1-to-3 checkpoint is unavailable offline, so no result here was obtained with the real one.
  * whole step -- ``oracle/train_step.py`` (fp32 or fp64) is the checker of ``tests/test_step_parity_gpu.py``.
"""
