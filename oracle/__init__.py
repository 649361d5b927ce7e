"""CPU oracle for the scHPF CAVI hot path -- TEST INFRASTRUCTURE ONLY.

hpf_numpy : numpy/scipy restatement (closest arithmetic to the reference:
            same scipy.special.digamma / gammaln symbols).
hpf_c     : ctypes binding of hpf_oracle.c (gcc + OpenMP), used where the
            numpy version would be too slow and as the "port" CPU baseline.

Nothing under schpf_b200/ imports this package.
"""
