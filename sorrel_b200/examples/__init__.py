"""Reference examples re-hosted on the B200 engine (treasurehunt, cleanup, iowa, tag).

Each sub-package keeps the reference's class names (`<Name>World`, `<Name>Env`, entity and agent classes) so a user
script only swaps its imports; `_common.py` holds what the four have in common on this side (config access, the
command-line runner).
"""
