"""
Minimal stand-in for the un-vendored ``fatiando==0.5`` dependency of the reference
(/root/reference/environment.yml:18).  TEST INFRASTRUCTURE ONLY: it lets the untouched
reference sources import, so that ``oracle/_ref`` can be built and the goldens replayed.
Nothing in the product package imports it.
"""
