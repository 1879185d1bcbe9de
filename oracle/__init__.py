"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the ORB-SLAM3 hot path.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  See orc_common.h for the pinning note
(the reference ships no tests; pinned against the reference's own sources
compiled unmodified into oracle/_ref/, and against cv2 4.13 underneath).
"""
from .oracle import *  # noqa: F401,F403
