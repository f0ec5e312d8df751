"""Host-side mirror of the reference's MPN circuit layer (Python, because no Rust toolchain exists in
this image — see DESIGN.md §0):

  cs.py        bellman's ConstraintSystem / LinearCombination / AllocatedNum / AllocatedBit / Boolean
               restated (un-vendored crate bellman 0.14.0), recording R1CS rows + witness values
  gadgets.py   /root/reference/src/zk/groth16/gadgets/{common,poseidon,merkle,eddsa}
  native.py    Fr helpers, Poseidon (/root/reference/src/zk/poseidon/mod.rs), JubJub + EdDSA
               (/root/reference/src/crypto/jubjub), sparse 4-ary Merkle state (/root/reference/src/zk/state/mod.rs)
  dw.py        deposit / withdraw transition builders, reveal gadget, DepositCircuit, WithdrawCircuit
               (/root/reference/src/mpn/{deposit,withdraw}.rs, circuits/{deposit,withdraw}_circuit.rs, gadgets/reveal)
  update.py    UpdateTransition builder (/root/reference/src/mpn/update.rs) and UpdateCircuit
               (/root/reference/src/mpn/circuits/update_circuit.rs)

The output of a circuit is (CSR R1CS, inputs, aux) for bazuka_b200.groth16.Prover — the GPU does the
proving; this layer is the witness / constraint generator the reference also runs on the CPU."""
