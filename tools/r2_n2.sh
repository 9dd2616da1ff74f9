#!/usr/bin/env bash
bash tools/r2_multi.sh 2
bash tools/r2_cfg.sh 2 cfg3_llama1m_512k_b4_n2 --model llama-3-8b-1m --ctx 524288 --batch 2
