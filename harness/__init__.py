"""TEST / BENCH HARNESS -- not product code, not shipped.

The real ``adaptaqc`` package needs qiskit, qiskit-aer and aqc_research, none of which can be installed in the build
image or on the GPU box.  So that the B200 backends can be DRIVEN here the way the reference drives its own backends,
this package restates, qiskit-free, the host-side parts of the reference that sit ABOVE the backend boundary:

  circuit.py     a QuantumCircuit-shaped container (``.data`` / ``.operation`` / ``.qubits``)
  compiler.py    ApproximateCompiler + AdaptCompiler loop      (adaptaqc/compilers/...)
  minimiser.py   Rotosolve / Rotoselect CostMinimiser           (adaptaqc/utils/cost_minimiser.py)
  measures.py    concurrence / EoF / negativity on 4x4 RDMs     (adaptaqc/utils/entanglement_measures.py)
  gradients.py   general_grad_of_pairs, the reference chain     (adaptaqc/utils/gradients.py)
  workloads.py   the BASELINE configs C3 / C4 / C5 as seeded generators

Only tests/, bench.py and __graft_entry__.smoke() import it.  The product package (adapt-aqc_b200/) never does
(tests/test_abi.py::test_product_never_imports_the_harness_or_the_oracle); a maintainer of the reference would ship
adapt-aqc_b200/ alone and keep using adaptaqc's own versions of everything in here.
"""
