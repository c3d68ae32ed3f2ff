"""oracle/ — TEST INFRASTRUCTURE ONLY (see vit3d_oracle.py header). Never imported by neurovit_b200."""
