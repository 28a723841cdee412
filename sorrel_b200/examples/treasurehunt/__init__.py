"""The treasurehunt example of the reference (sorrel/examples/treasurehunt) on the B200 engine."""
