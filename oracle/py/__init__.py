"""ORACLE — CPU restatement of the reference's algorithms on Python big integers.
Test infrastructure only: imported by tests/, __graft_entry__.smoke() and nothing else."""
