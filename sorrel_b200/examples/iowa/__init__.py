"""The iowa (gambling task) example of the reference (sorrel/examples/iowa) on the B200 engine."""
