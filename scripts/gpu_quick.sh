#!/bin/bash
AB_BATCH=1 timeout 600 python scripts/ab_step.py pdl=caco_set_pdl:1 nopdl=caco_set_pdl:0 2>&1 | tail -1
AB_BATCH=8 timeout 600 python scripts/ab_step.py pdl=caco_set_pdl:1 nopdl=caco_set_pdl:0 2>&1 | tail -1
