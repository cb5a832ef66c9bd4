"""TEST INFRASTRUCTURE ONLY: CPU oracle for the moolib hot paths (see moolib_oracle.c / oracle.py).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from .oracle import *  # noqa: F401,F403
