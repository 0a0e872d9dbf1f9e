"""Baseline arms of bench.py (not product code): the reference's loop statements in stock PyTorch on the same GPU."""
