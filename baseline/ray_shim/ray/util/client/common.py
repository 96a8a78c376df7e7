class ClientActorHandle:
    pass
