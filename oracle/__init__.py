"""Parity oracle for the MILAN describe-neurons path. TEST INFRASTRUCTURE ONLY — never imported by the product."""
