# Minimal stand-in for the `timm` package (absent from this image; no network).
# Test infrastructure only: lets oracle/make_golden.py import the reference's
# network/model_parts.py, which needs timm.layers.{DropPath,to_2tuple,trunc_normal_}.
