"""Test-infrastructure stand-in for the undeclared `torchmeta` dependency of the
reference (modules.py:4).  Only used by oracle/gen_golden.py, in the container
where /root/reference exists, to import the reference UNCHANGED.  Container
classes only - no arithmetic lives here (SplitFCBlock never passes `params`,
modules.py:247,260,266,316,353)."""
