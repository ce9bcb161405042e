"""Host-side mirror of the CUDA-L2 evaluation harness for --device_type b200 (see common.py)."""
