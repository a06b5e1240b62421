"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/oracle_orb.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package never does.
"""
from .binding import *  # noqa
