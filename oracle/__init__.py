"""CPU ORACLE package (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
`oracle.lib`  : ctypes binding of the C++ restatement (oracle/ckc_oracle*.cpp -> oracle/_build/libckc_oracle.so)
`oracle.refshim` : driver of the genuine reference code inside the reference's prebuilt tests/trbspcs (oracle/_ref/)
"""
