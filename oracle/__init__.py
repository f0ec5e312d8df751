"""ORACLE package (test infrastructure only). See oracle/README.md."""
