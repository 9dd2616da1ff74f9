#!/usr/bin/env bash
bash tools/r2_multi2.sh 4
bash tools/r2_cfg.sh 4 cfg3_llama1m_512k_b4_n4 --model llama-3-8b-1m --ctx 524288 --batch 1
