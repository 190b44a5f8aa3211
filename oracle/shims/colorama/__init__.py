class _S:
    def __getattr__(self, k):
        return ""
Fore = _S(); Style = _S(); Back = _S()
