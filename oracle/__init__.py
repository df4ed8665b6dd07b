"""Parity oracle for the CUDA rasterizer -- TEST INFRASTRUCTURE ONLY.

  gs_oracle.c / oracle.py   CPU restatement of the reference rasterizer (C + numpy front-end)
  build.py                  gcc recipe for libgs_oracle.so
  build_ref.py              nvcc recipe that compiles the UNMODIFIED reference into oracle/_ref/
  ref_api.py                loads oracle/_ref (the reference's own Python API) under a private name

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this package.  Nothing under generativedensification_b200/ or
diff_gaussian_rasterization/ does.
"""
