"""TEST INFRASTRUCTURE ONLY -- stand-in for the stdlib `turtle` (needs tkinter, absent here).

`/root/reference/loss_trainer.py:1` does `from turtle import pd` (a stray import, never used)."""
pd = None
