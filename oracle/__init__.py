"""CPU oracle for the GANet guided-aggregation hot path.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(ganet_b200/) never imports this.

  oracle.api      numpy front-end over libganet_oracle.so (our C restatement)
  oracle.ref_cpu  numpy front-end over oracle/_ref/libganet_ref_cpu.so (the
                  reference's own kernel bodies compiled for the host)
  oracle.ref_gpu  loader for oracle/_ref/GANet*.so (the unmodified reference
                  CUDA extension; needs a GPU)
"""
