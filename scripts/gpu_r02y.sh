#!/bin/bash
bash scripts/sanitize.sh
