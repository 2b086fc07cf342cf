"""TEST INFRASTRUCTURE ONLY: CPU restatements of the reference's arithmetic (torch-functional graph, plain-C primitives, numpy
metrics, loader arithmetic).  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm -- never by
heatnet_pub_b200 (tests/test_host.py::test_product_path_never_touches_the_oracle)."""
