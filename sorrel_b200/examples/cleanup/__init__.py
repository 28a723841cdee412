"""The cleanup example of the reference (sorrel/examples/cleanup) on the B200 engine."""
