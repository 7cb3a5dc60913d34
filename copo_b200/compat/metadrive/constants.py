DEFAULT_AGENT = "default_agent"          # metadrive.constants.DEFAULT_AGENT (used at torch_copo/algo_ippo.py:50)
