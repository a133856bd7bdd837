"""CPU oracle of the UNet -> WS hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (ws_unet_b200) never does and fails loudly without its CUDA library.

Pinning status (DESIGN.md section "Oracle"):
  * ws_oracle (numpy/C WS arithmetic, linear filters): PINNED against the reference's own outputs -
    tests/golden/ws_golden.npz holds beta_hat / predictions produced by importing the unmodified reference
    (src/ws/estimate.py::attack, src/filters/evaluate.py::infere_single) on crops of its shipped images, and
    tests/test_oracle_reference.py re-checks results/prediction/filters.csv when /root/reference is present.
  * unet_oracle (C forward pass): PINNED against the reference's own torch module
    (src/unet/model/unet.py::UNet, loaded with deterministic numpy weights) via tests/golden/unet_golden.npz.
"""
