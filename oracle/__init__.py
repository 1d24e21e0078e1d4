"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's hot path (see oracle/oracle.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package."""
