"""CPU oracle for the biased-Langevin VES hot path of apallath/pyves.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``pyves_b200/`` imports it, and the product path
raises when the CUDA library is missing instead of falling back to this code.

What is restated here (FP64 numpy; every function cites the reference file:line it follows):

* ``legendre``   -- ``ves/basis.py:86-156`` (Bonnet recursion, scaling, clamp), plus the 2D
                    tensor-product extension the north-star asks for (no reference symbol).
* ``mlp``        -- ``ves/basis.py:184-232`` (``NNBasis1D`` layer structure and scaling).
* ``pathcv``     -- ``ves/basis.py:509-548`` (path collective variable basis).
* ``potentials`` -- ``ves/potentials.py`` numpy twins + hand-derived gradients.
* ``langevin``   -- the update ``openmm.LangevinIntegrator.step`` performs at
                    ``ves/langevin_dynamics.py:188`` (OpenMM is an un-vendored dependency,
                    ``requirements.txt:3`` ``OpenMM>=7.6.0``; algorithm restated from the
                    published ``ReferenceStochasticDynamics`` scheme).
* ``philox``     -- Philox4x32-10 (Salmon et al., SC'11) + the Box-Muller mapping the CUDA
                    kernels use, so RNG-driven runs can be compared walker by walker.
* ``ves``        -- ``ves/bias.py:206-253`` (reference-literal loss) and the textbook VES
                    gradient / Hessian / averaged SGD / well-tempered target (literature only).
* ``fes``        -- ``ves/visualization.py:185-201,221-237,475-478`` (analytic and histogram FES).

Pin status
----------
* PINNED against the unmodified reference (fixtures in ``tests/golden/`` made by
  ``tests/golden/make_golden.py``, which imports ``/root/reference/ves`` in the build
  container): Legendre 1D energies/forces, NN energies/forces, path-CV energies/forces,
  the ``update()`` loss/gradient/weights for SGD/Adam/RMSprop, all potentials, the target,
  the analytic FES projections; plus the scipy Legendre known answers of
  ``tests/tests_unit/test_basis.py`` and the Random123 Philox known-answer vectors.
* PARITY UNPINNED (no reference implementation exists or can run here): the Langevin
  integrator (OpenMM absent from the image and from ``/root/reference``), the 2D
  tensor-product basis, the ensemble/textbook VES optimiser, the well-tempered target and
  the basis covariance.  For those the oracle is the literature restatement and the analytic
  free-energy surface, and the tests say so.
"""
