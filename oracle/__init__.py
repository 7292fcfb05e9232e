"""CPU oracle for the exact expectation-value path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``ml_qem_b200``) never
does, and fails loudly when its CUDA library is missing.

What it restates
----------------
The reference (qiskit-community/ml-qem, python package ``blackwater``) contains no simulator
arithmetic of its own: ``blackwater/data/utils.py:418-431`` (create_estimator_meas_data) and
``:434-444`` call ``qiskit_aer.primitives.Estimator`` / ``AerSimulator.from_backend`` and
``docs/tutorials/h13_ising_data_gen_tomo.ipynb:811`` calls ``qiskit.primitives.Estimator``.
Those live in un-vendored third-party packages pinned by ``requirements.txt:3-5``
(qiskit-aer>=0.11.0,<=0.13.3; qiskit==0.43.2 => qiskit-terra 0.24.1).  Neither is
installable here (no network), so this package restates their published algorithms:

* ``gates``        Qiskit standard-gate matrices, little-endian (names: utils.py:19-49)
* ``noise_model``  ``NoiseModel.from_backend`` device model (depolarizing o thermal relaxation)
                   as used at utils.py:427, plus the in-tree editors
                   docs/tutorials/noise_utils.py:36-144 and docs/tutorials/mbd_utils.py:95-137
* ``dm``           Aer ``density_matrix`` method: column-stacked vec(rho) in complex128, every
                   gate applied as conj(U) (x) U and every attached error as one superoperator
                   AFTER its gate; ``Tr(rho P)`` as in docs/tutorials/vqe_rf.py:57-83
* ``sv``           statevector evolution + <psi|P|psi> (terra ``Estimator``, shots=None)

Parity status: the reference's own tests pin no simulator value, so parity is anchored on the
known answers the reference stores in notebooks/data (``tests/golden/*.json``, produced by
``tests/golden/make_golden.py``): Aer's noise-model dump for FakeLima (mixture probabilities
to full precision, Kraus operators to 8 digits), mean average-gate infidelities of the Aer
noise models (FakeLima / FakeBelem, incoherent and coherent-CX variants, 16 digits), 10k-shot
ideal/noisy values of FakeLima-transpiled circuits (statistical), and the H2 Hamiltonians'
FCI energies.  ``tests/test_oracle_golden.py`` checks all of them.  Real Aer could not be
executed, so end-to-end parity "vs Aer" is transitive: oracle == Aer known answers, engine ==
oracle to 1e-10.
"""
