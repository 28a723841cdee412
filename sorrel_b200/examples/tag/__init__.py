"""The tag example of the reference (sorrel/examples/tag) on the B200 engine."""
