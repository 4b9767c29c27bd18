"""CPU oracle for the JWave wavelet hot path - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (jwave_b200/) must never import it.
"""
