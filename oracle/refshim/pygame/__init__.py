"""empty stand-in (CybORG/env.py imports pygame for a disabled GUI)"""
