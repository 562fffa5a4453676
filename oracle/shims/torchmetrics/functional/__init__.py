"""TEST INFRASTRUCTURE ONLY: stand-in so that the reference's train.py / test.py import here (torchmetrics is not
installed; its text metrics are not on the classification path)."""
