"""Stand-ins for third-party packages the reference's host-side code imports but this stack does not ship.  They are part of
the launcher's boundary (relightable_nr_b200/run.py), not of the hot path."""
