"""Placeholder for ``pyquil.api`` so ``import pyquil.api`` does not fail; no cloud access here."""
